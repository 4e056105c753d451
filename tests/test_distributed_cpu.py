"""Data-parallel plumbing with world_size 2 on the gloo backend (CPU): one flat gradient bucket per optimizer is
all-reduced, and every rank ends up with the parameters a single process would get from the mean gradient.

The CUDA kernels cannot run here; the test substitutes torch restatements of the three launches (`_gather_grads`,
`_grad_norm`, `_kernel_step`) — test infrastructure only — so that bucket construction, static membership, the
all-reduce and the 1/world scaling are exercised exactly as on the GPU."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _torch_kernel_step(self, b, group, step, max_norm, inv_scale):
    g = b.flat_g * inv_scale
    total = g.norm()
    coef = min(1.0, max_norm / (float(total) + 1e-6)) if max_norm > 0 else 1.0
    g = g * coef
    beta1, beta2 = group["betas"]
    b.flat_p.mul_(1 - group["lr"] * group["weight_decay"])
    b.m.mul_(beta1).add_(g, alpha=1 - beta1)
    b.v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    b.flat_p.addcdiv_(b.m, (b.v.sqrt() / bc2 ** 0.5).add_(group["eps"]), value=-group["lr"] / bc1)


def _torch_gather(self, b, fused_norm):
    for p in b.params:
        b.view_of(b.flat_g, p).copy_(p.grad if p.grad is not None else torch.zeros_like(p))
    if fused_norm:
        b.stats[0] += (b.flat_g ** 2).sum()


def _torch_grad_norm(self, b):
    b.stats[0] += (b.flat_g ** 2).sum()


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from optispeech_b200.optim import FlatAdamW

    FlatAdamW._kernel_step = _torch_kernel_step
    FlatAdamW._gather_grads = _torch_gather
    FlatAdamW._grad_norm = _torch_grad_norm
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(17, 5)), torch.nn.Parameter(torch.randn(33)), torch.nn.Parameter(torch.randn(4, 4))]
    frozen = torch.nn.Parameter(torch.randn(6))  # never receives a gradient on any rank: static membership excludes it
    opt = FlatAdamW([{"params": params + [frozen]}], lr=1e-2, betas=(0.8, 0.99), weight_decay=1e-2, max_grad_norm=10.0, loss_scale=8.0,
                    world_size=world)
    for it in range(3):
        opt.zero_grad()
        g = torch.Generator().manual_seed(100 * it + rank)
        for p in params:
            gr = torch.randn(p.shape, generator=g) * 8.0  # carries the loss scale
            if p.grad is None:
                p.grad = gr
            else:
                p.grad.add_(gr)
        opt.step()
    torch.save([p.detach().clone() for p in params] + [frozen.detach().clone()], os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_bucket_allreduce_matches_single_process(tmp_path):
    port = 29500 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    for a, b in zip(r0, r1):
        assert torch.equal(a, b), "ranks diverged"
    # single-process reference: torch.optim.AdamW on the mean gradient with clip_grad_norm_(10)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(17, 5)), torch.nn.Parameter(torch.randn(33)), torch.nn.Parameter(torch.randn(4, 4))]
    frozen = torch.randn(6)
    opt = torch.optim.AdamW(params, lr=1e-2, betas=(0.8, 0.99), weight_decay=1e-2)
    for it in range(3):
        gens = [torch.Generator().manual_seed(100 * it + r) for r in range(2)]
        for p in params:
            p.grad = sum(torch.randn(p.shape, generator=g) for g in gens) / 2
        torch.nn.utils.clip_grad_norm_(params, 10.0)
        opt.step()
    for a, b in zip(r0[:3], params):
        assert torch.allclose(a, b.detach(), rtol=1e-5, atol=1e-6)
    assert torch.equal(r0[3], frozen), "parameter without gradient must stay untouched (no weight decay)"


# --------------------------------------------------------------------------------------------------
# synthesis: replicas only — utterances are sharded by length, results gathered on the host (utils/sharding.py)
# --------------------------------------------------------------------------------------------------
class _EchoSynth(torch.nn.Module):
    """Stand-in for OptiSpeechGenerator.synthesise on the CPU: two samples per phoneme id value, so that the gathered result
    identifies the utterance it came from (the kernels cannot run here; the sharding / batching / gather logic is what is tested)."""

    def __init__(self):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1))
        self.calls = []

    def synthesise(self, x, x_lengths, **kw):
        self.calls.append(tuple(int(v) for v in x_lengths))
        assert x.shape[1] == int(x_lengths.max()), "a batch is cut to its longest utterance"
        wav_lengths = x_lengths * 2
        wav = torch.zeros(x.shape[0], int(wav_lengths.max()))
        for b in range(x.shape[0]):
            n = int(x_lengths[b])
            wav[b, : 2 * n] = x[b, :n].float().repeat_interleave(2)
        return {"wav": wav, "wav_lengths": wav_lengths, "durations": torch.full_like(x, 2)}


def _synth_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from optispeech_b200.utils.sharding import gather_outputs, synthesise_sharded

    g = torch.Generator().manual_seed(3)
    N = 11
    lens = torch.randint(5, 40, (N,), generator=g)
    x = torch.zeros(N, int(lens.max()), dtype=torch.int64)
    for i in range(N):
        x[i, : lens[i]] = torch.randint(1, 159, (int(lens[i]),), generator=g)
    model = _EchoSynth()
    local = synthesise_sharded(model, x, lens, rank, world, max_batch=3)
    merged = gather_outputs(local, world)
    torch.save({"local": sorted(local), "merged": {k: v["wav"] for k, v in merged.items()}, "calls": model.calls,
                "expect": {i: x[i, : lens[i]].float().repeat_interleave(2) for i in range(N)}}, os.path.join(out_dir, f"synth{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_synthesis_shards_by_length_and_gathers(tmp_path):
    from optispeech_b200.utils.sharding import shard_by_length

    shards = shard_by_length([10, 50, 30, 30, 5, 70, 20], 3)
    assert sorted(i for s in shards for i in s) == list(range(7))                  # a partition
    assert [max(len(s) for s in shards) - min(len(s) for s in shards)] == [1]
    loads = [sum([10, 50, 30, 30, 5, 70, 20][i] for i in s) for s in shards]
    assert max(loads) - min(loads) <= 70                                           # within one utterance of each other
    assert shard_by_length([3, 1, 2], 1) == [[0, 2, 1]]                            # longest first
    with pytest.raises(ValueError):
        shard_by_length([1], 0)

    port = 29900 + (os.getpid() % 90)
    mp.spawn(_synth_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "synth0.pt"), torch.load(tmp_path / "synth1.pt")
    assert not set(r0["local"]) & set(r1["local"]) and sorted(r0["local"] + r1["local"]) == list(range(11))
    for r in (r0, r1):                                                             # every rank holds every waveform after the gather
        assert sorted(r["merged"]) == list(range(11))
        for i, w in r["merged"].items():
            assert torch.equal(w, r["expect"][i]), i
        assert all(len(c) <= 3 for c in r["calls"])                                # max_batch honoured
        assert all(list(c) == sorted(c, reverse=True) for c in r["calls"])         # similar lengths share a batch, longest first
