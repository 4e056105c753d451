"""Data parallelism on two GPUs over NCCL (needs >= 2 devices; skipped otherwise): the flat gradient bucket of the training
step is all-reduced, eagerly and inside the captured CUDA graph, every rank ends with bit-identical parameters, and the
reduced bucket is the mean of the ranks' local gradients (reference: Lightning DDP around base_lightning_module.py:86-130)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir, graph):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from functools import partial

    from test_public_surface_gpu import ModelSpec, _fresh_model, _small_batch

    spec = ModelSpec()
    batch = _small_batch(spec, 3, 48, 200, seed=7, dev=dev)
    if rank == 1:   # same shapes and lengths, different content
        batch = dict(batch)
        batch["mel"], batch["pitches"], batch["energies"] = batch["mel"] * 0.7, -batch["pitches"], batch["energies"] * 0.5
    out = {}
    # (a) lr = 0: the bucket after the step is the all-reduced mean gradient (times the loss scale)
    model = _fresh_model(spec, dev)
    model.hparams.optimizer = partial(torch.optim.AdamW, lr=0.0, betas=[0.8, 0.99], weight_decay=0.0)
    model.cuda_graph = graph
    for i in range(6):
        model.training_step(batch, i)
    torch.cuda.synchronize()
    bucket = model.optimizers()[0].buckets()[0]
    reduced = bucket.flat_g.clone()
    if model._graphed is not None:
        model._graphed.release()
    # local gradient of this rank through plain autograd, in bucket order
    for p in model.generator.parameters():
        p.grad = None
    o = model._process_batch(batch)
    (o["loss"] * model.loss_scale).backward()
    from optispeech_b200 import ops
    ops.join_grad_streams()
    local = torch.zeros_like(reduced)          # members sit at 16-byte aligned offsets of the bucket
    for p, o in zip(bucket.params, bucket.offsets):
        if p.grad is not None:
            local[o:o + p.numel()] = p.grad.reshape(-1).float()
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    sum_local = sum(gathered)                  # the all-reduce SUMS; 1/world is folded into the optimizer's unscale factor
    out["mean_rel"] = float((reduced - sum_local).norm() / sum_local.norm())
    out["own_rel"] = float((reduced - local).norm() / local.norm())
    # (b) real steps: ranks stay bit-identical
    model2 = _fresh_model(spec, dev)
    model2.cuda_graph = graph
    for i in range(6):
        model2.training_step(batch, i)
    torch.cuda.synchronize()
    out["params"] = {n: p.detach().cpu().clone() for n, p in model2.generator.named_parameters()}
    if model2._graphed is not None:
        model2._graphed.release()
    # (c) GAN phase: two optimizers, two all-reduces per step — the discriminator turn's is issued from the turn's own stream
    # next to the generator's backward pass (BaseModule.OVERLAP_TURNS); both ranks must still end bit-identical, generator
    # and discriminators alike
    torch.manual_seed(4321)                    # same discriminator initialisation on both ranks
    model3 = _fresh_model(spec, dev)
    model3.train_args.pretraining_steps = 0
    model3.cuda_graph = graph
    for i in range(6):
        model3.training_step(batch, i)
    torch.cuda.synchronize()
    assert "total_loss/discriminator" in model3.logged
    out["gan_params"] = {n: p.detach().cpu().clone() for n, p in model3.named_parameters()}
    out["gan_losses"] = (float(model3.logged["total_loss/generator"]), float(model3.logged["total_loss/discriminator"]))
    if model3._graphed is not None:
        model3._graphed.release()
    torch.save(out, os.path.join(out_dir, f"rank{rank}_{int(graph)}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("graph", [False, True])
def test_two_rank_training_step_over_nccl(tmp_path, graph):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29700 + (os.getpid() % 200) + int(graph)
    mp.spawn(_worker, args=(2, port, str(tmp_path), graph), nprocs=2, join=True)
    r0 = torch.load(tmp_path / f"rank0_{int(graph)}.pt")
    r1 = torch.load(tmp_path / f"rank1_{int(graph)}.pt")
    print(f"graph={graph}: |reduced - sum(local)| / |sum(local)| = {r0['mean_rel']:.3e}; "
          f"|reduced - own local| / |own local| = {r0['own_rel']:.3e} (rank 0), {r1['own_rel']:.3e} (rank 1)")
    diverged = [(n, float((a - r1["params"][n]).abs().max())) for n, a in r0["params"].items() if not torch.equal(a, r1["params"][n])]
    assert not diverged, f"ranks diverged on {len(diverged)} of {len(r0['params'])} tensors, e.g. {diverged[:6]}"
    gan_diverged = [(n, float((a - r1["gan_params"][n]).abs().max())) for n, a in r0["gan_params"].items() if not torch.equal(a, r1["gan_params"][n])]
    print(f"graph={graph}: GAN phase losses rank 0 {r0['gan_losses']}, rank 1 {r1['gan_losses']}; {len(r0['gan_params'])} tensors compared")
    assert not gan_diverged, f"GAN phase: ranks diverged on {len(gan_diverged)} of {len(r0['gan_params'])} tensors, e.g. {gan_diverged[:6]}"
    assert all(v == v and abs(v) < 1e6 for v in r0["gan_losses"] + r1["gan_losses"])
    # run-to-run floor of one gradient evaluation is a few percent on the predictors (fp32 atomics + fp16 roundings); a
    # rank's own gradient differs from the sum by O(1)
    assert r0["mean_rel"] <= 5e-2 and r1["mean_rel"] <= 5e-2
    assert r0["own_rel"] >= 4 * r0["mean_rel"] and r1["own_rel"] >= 4 * r1["mean_rel"]
