"""ctypes binding of libosb200.so (declared in include/osb200.h).

The product path has no fallback: importing this module without a built library, or calling
into it on a machine without a B200, raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = _PKG_DIR / "libosb200.so"

OSB_OK = 0

# osb_epilogue
EPI_BIAS, EPI_GELU, EPI_RESID, EPI_RELU_LN, EPI_BIAS_LN, EPI_RELU, EPI_GELU_BWD, EPI_LN_BWD, EPI_RELU_LN_BWD, EPI_RELU_BWD, EPI_ATTN_LOGP, EPI_AXPY = range(12)
# flags
FLAG_CLIP, FLAG_KEEPMASK, FLAG_OUT_H16, FLAG_SAVE_PRE, FLAG_DOT, FLAG_SPLIT_IN, FLAG_SPLIT_OUT, FLAG_RELU, FLAG_NO_F32, FLAG_COLSUM, FLAG_W_MN, FLAG_TAP_REVERSE, FLAG_LRELU = 1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096


class GemmDesc(C.Structure):
    """Mirror of `osb_gemm_desc` (include/osb200.h)."""

    _fields_ = [
        ("a", C.c_void_p),
        ("w", C.c_void_p),
        ("lda", C.c_int64),
        ("ldw", C.c_int64),
        ("B", C.c_int32),
        ("T", C.c_int32),
        ("N", C.c_int32),
        ("K", C.c_int32),
        ("taps", C.c_int32),
        ("pad", C.c_int32),
        ("epi", C.c_int32),
        ("flags", C.c_int32),
        ("out", C.c_void_p),
        ("aux_h16", C.c_void_p),
        ("ldo", C.c_int64),
        ("bias", C.c_void_p),
        ("resid", C.c_void_p),
        ("gamma", C.c_void_p),
        ("row_scale", C.c_void_p),
        ("pad_mask", C.c_void_p),
        ("ln_w", C.c_void_p),
        ("ln_b", C.c_void_p),
        ("ln_eps", C.c_float),
        ("dot_w", C.c_void_p),
        ("dot_b", C.c_void_p),
        ("out_dot", C.c_void_p),
        ("aux_in_h16", C.c_void_p),
        ("row_stat", C.c_void_p),
        ("dropout_p", C.c_float),
        ("dropout_seed", C.c_uint64),
        ("dropout_seed_dev", C.c_void_p),
        ("w_batched", C.c_int32),
        ("col_len", C.c_void_p),
        ("out_colsum", C.c_void_p),
        ("row_stride", C.c_int32),
        ("T_in", C.c_int32),
        ("lrelu_slope", C.c_float),
        ("seq_pitch", C.c_int32),
        ("seq_valid", C.c_int32),
    ]


class PackJob(C.Structure):
    """Mirror of `osb_pack_job` (include/osb200.h)."""

    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("col_scale", C.c_void_p), ("aux", C.c_void_p), ("first_elem", C.c_int64), ("kind", C.c_int32),
                ("rows", C.c_int32), ("cols", C.c_int32), ("k", C.c_int32), ("dst_cols", C.c_int32), ("reserved", C.c_int32)]


class OsbError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """Load libosb200.so, failing loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise OsbError(
            f"{LIB_PATH} is missing: run `python -m optispeech_b200.build` (or __graft_entry__.build()). "
            "There is no CPU/PyTorch fallback for the B200 hot path."
        )
    lib = C.CDLL(str(LIB_PATH))
    lib.osb_version.restype = C.c_int
    lib.osb_strerror.restype = C.c_char_p
    lib.osb_strerror.argtypes = [C.c_int]
    lib.osb_launch_count.restype = C.c_ulonglong
    lib.osb_check_device.argtypes = [C.c_int]
    lib.osb_gemm.argtypes = [C.POINTER(GemmDesc), C.c_void_p]
    lib.osb_gemm_wgrad.argtypes = [
        C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
        C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
    ]
    P, I32, I64, F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    sigs = {
        "osb_convnext_block_fwd": [P, P, P, P, P, P, P, P, P, P, P, I32, I32, I32, I32, F, P],
        "osb_embed_text": [P, P, P, P, P, I32, I32, I32, I32, P],
        "osb_dwconv_ln": [P, P, P, P, P, I32, I32, I32, F, I32, P],
        "osb_layernorm": [P, P, P, P, P, I64, I32, F, I32, P],
        "osb_relu_layernorm": [P, P, P, P, I64, I32, F, I32, P, P, P, P, P],
        "osb_variance_embed": [P, P, P, P, P, P, P, P, I32, I32, I32, I32, I32, P],
        "osb_durations": [P, P, P, P, I32, I32, F, F, P],
        "osb_centres": [P, I32, P, P, I32, I32, P],
        "osb_gaussian_upsample": [P, P, P, P, P, P, I32, I32, I32, I32, F, P],
        "osb_gaussian_upsample_window": [P, P, P, P, P, P, P, I32, I32, I32, I32, I32, I32, F, P],
        "osb_expand_gather": [P, P, P, P, I32, I32, I32, I32, P],
        "osb_pack_h16": [P, I64, I64, P, P, P, I64, I32, I64, I32, P],
        "osb_pack_conv_h16": [P, P, P, I32, I32, I32, I32, I32, P],
        "osb_resid_bwd_prep": [P, P, P, P, P, P, P, P, I64, I32, I32, P],
        "osb_colsum_h16": [P, P, I64, I32, P],
        "osb_ln_fold_bwd": [P, P, P, P, P, P, P, I32, I32, P],
        "osb_dwconv_bwd": [P, P, P, P, P, P, P, P, I32, I32, I32, P],
        "osb_layernorm_bwd": [P, P, P, P, P, P, I64, I32, F, P],
        "osb_predictor_tail_bwd": [P, P, P, P, P, P, P, P, P, P, P, I64, I32, F, F, C.c_uint64, P, P],
        "osb_ln_param_grad": [P, P, P, P, P, I64, I32, F, P],
        "osb_variance_embed_bwd": [P, P, P, P, P, P, P, I32, I32, I32, I32, P],
        "osb_embed_text_bwd": [P, P, P, P, P, I32, I32, I32, I32, I32, P],
        "osb_mas": [P, P, P, P, P, I32, I32, I32, P],
        "osb_gemm_wgrad_batched": [P, I64, P, I64, P, I32, I32, I32, I32, P],
        "osb_gemm_wgrad_strided": [P, I64, P, I64, P, I32, I32, I32, I32, I32, I32, I32, I32, P],
        "osb_rownorm_sq": [P, P, I64, I32, P],
        "osb_forward_sum": [P, P, P, F, P, P, P, I32, I32, I32, P],
        "osb_beta_binomial_prior": [P, I64, P, P, P, I32, I32, I32, P],
        "osb_attn_bwd_prep": [P, P, P, P, P, P, P, P, P, I32, I32, I32, I32, P],
        "osb_transpose_pack_h16": [P, P, I32, I32, I32, I32, P],
        "osb_scale_rows": [P, P, P, I64, I32, F, P],
        "osb_stft_loss": [P, P, P, I32, I32, I32, I32, I32, F, P, P, P, P],
        "osb_mel_loss": [P, P, P, P, P, P, P, P, I32, I32, I32, I32, I32, I32, F, P, P, P, P],
        "osb_grad_sumsq": [P, I64, P, P],
        "osb_grad_gather": [P, P, I64, P, P, P],
        "osb_adamw_step": [P, P, P, P, I64, P, F, F, F, F, F, I64, F, F, P],
        "osb_adamw_step_dev": [P, P, P, P, I64, P, P, F, F, F, F, F, F, P],
        "osb_average_by_duration": [P, P, P, P, P, I32, I32, I32, P],
        "osb_pack_multi": [P, I32, I64, P],
        "osb_align_loss_fold": [P, P, P, P, P, P, P, I32, I32, I32, P],
        "osb_fs2_losses": [P, P, P, P, P, P, P, P, P, P, P, I32, I32, P],
        "osb_mha_fwd": [P, P, P, I64, P, P, I64, I64, P, P, I32, I32, I32, I32, F, F, C.c_uint64, P, P],
        "osb_mha_bwd": [P, P, P, I64, P, P, I64, P, I64, P, P, P, I64, P, P, I64, I32, I32, I32, I32, I32, F, F, C.c_uint64, P, P],
        "osb_mha_pack_heads": [P, P, I64, I32, I32, I32, P],
        "osb_dropout_pack_h16": [P, P, I64, I32, F, C.c_uint64, P, P],
        "osb_add_posenc": [P, P, P, P, I32, I32, I32, F, C.c_uint64, P, P],
        "osb_convnext_block_fwd_train": [P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, I32, I32, I32, I32, F, P],
        "osb_convnext_block_bwd": [P, P, P, P, P, P, P, P, P, P, I32, I32, I32, I32, P],
        "osb_ln_dwconv_bwd": [P, I32, P, P, P, P, P, P, P, P, I32, I32, I32, P],
        "osb_convnext_block_bwd_parts": [I32, I32, I32],
        "osb_resid_param_grad": [P, P, P, P, P, P, P, P, I64, I32, I32, P],
        "osb_sequence_mask": [P, P, P, I32, I32, P],
        "osb_segment_starts": [P, P, P, I32, I32, I32, P],
        "osb_gather_segments": [P, P, P, I32, I64, I32, I32, I32, P],
        "osb_mrd_first_fwd": [P, P, P, P, I32, I32, I32, I32, I32, I32, F, P],
        "osb_mrd_first_bwd": [P, P, P, P, P, P, I32, I32, I32, I32, I32, I32, F, P],
        "osb_spec_im2col_h16": [P, P, I32, I32, I32, I32, I32, I32, P],
        "osb_spec_col2im": [P, P, I32, I32, I32, I32, I32, I32, F, P],
        "osb_wim2col_h16": [P, P, I32, I32, I32, I32, I32, I32, I32, I32, P],
        "osb_wcol2im_h16": [P, P, I32, I32, I32, I32, I32, I32, I32, I32, P],
        "osb_mrd_post_fwd": [P, P, P, P, I32, I32, I32, I32, P],
        "osb_mrd_post_bwd": [P, P, P, P, P, P, I32, I32, I32, I32, F, P],
        "osb_mel_energy": [P, P, P, P, P, P, P, P, I32, I32, I32, I32, I32, I32, I32, F, F, P],
        "osb_mpd_first_fwd": [P, P, P, P, I32, I32, I32, I32, I32, I32, I32, F, P],
        "osb_mpd_first_bwd": [P, P, P, P, P, P, I32, I32, I32, I32, I32, I32, I32, F, P],
        "osb_mpd_post_fwd": [P, P, P, P, I32, I32, I32, I32, I32, P],
        "osb_mpd_post_bwd": [P, P, P, P, P, P, I32, I32, I32, I32, I32, F, P],
        "osb_lrelu_bwd_h16": [P, P, P, I64, I32, I32, I32, F, P, F, P],
        "osb_col2im_h16": [P, P, I64, I64, I32, I32, I32, I32, I32, P],
        "osb_l1_pair_fwd": [P, P, P, I64, P],
        "osb_l1_pair_bwd": [P, P, P, F, P, I64, P],
    }
    for name, argtypes in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.osb_gemm.restype = C.c_int
    lib.osb_gemm_wgrad.restype = C.c_int
    _lib = lib
    return lib


def exported_symbols() -> list[str]:
    """Names every entry point include/osb200.h declares (checked by the CPU test-suite)."""
    import re

    header = (_PKG_DIR.parent / "include" / "osb200.h").read_text()
    return sorted(set(re.findall(r"\b(osb_[a-z0-9_]+)\s*\(", header)))


def check(status: int, what: str) -> None:
    if status == OSB_OK:
        return
    lib = load()
    if status < 0:
        raise OsbError(f"{what}: {lib.osb_strerror(status).decode()} (osb_status {status})")
    raise OsbError(f"{what}: CUDA error {status}")


def launch_count() -> int:
    return int(load().osb_launch_count())


# --------------------------------------------------------------------------------------------------
# optional per-launch timing (bench.py's roofline pass): CUDA events on the launching stream around every entry point
# --------------------------------------------------------------------------------------------------
class LaunchProfiler:
    """Wraps the library's entry points so that every call is bracketed by CUDA events recorded on the current stream.
    Used for ONE separate profiling pass after the timed region (event recording perturbs timing)."""

    def __init__(self):
        self.records = []  # (name, key, flops, start_event, end_event)
        self.calls = []    # (raw entry point, ctypes argument tuple) of every record, for replay_amortized()
        self.amortized = {}  # record index -> device ms per launch from replay_amortized()
        self._saved = {}

    def __enter__(self):
        import torch

        lib = load()
        for name in exported_symbols():
            if name in ("osb_version", "osb_strerror", "osb_launch_count", "osb_check_device"):
                continue
            fn = getattr(lib, name)
            self._saved[name] = fn

            def make(fn=fn, name=name):
                def wrapped(*args):
                    key, flops = name, 0.0
                    if name == "osb_gemm":
                        d = args[0]._obj
                        split = 3 if (d.flags & FLAG_SPLIT_IN) else 1
                        key = f"osb_gemm[epi={d.epi},rows={d.B * d.T},N={d.N},K={d.K},taps={d.taps},passes={split}]"
                        flops = 2.0 * d.B * d.T * d.N * d.K * d.taps
                    elif name == "osb_convnext_block_fwd":
                        B, T, Cc, I = args[11], args[12], args[13], args[14]
                        key = f"osb_convnext_block_fwd[rows={B * T},C={Cc},I={I}]"
                        flops = 2.0 * B * T * (2 * Cc * I + 7 * Cc)
                    elif name == "osb_convnext_block_fwd_train":
                        B, T, Cc, I = args[15], args[16], args[17], args[18]
                        key = f"osb_convnext_block_fwd_train[rows={B * T},C={Cc},I={I}]"
                        flops = 2.0 * B * T * (2 * Cc * I + 7 * Cc)
                    elif name == "osb_convnext_block_bwd":
                        B, T, Cc, I = args[10], args[11], args[12], args[13]
                        key = f"osb_convnext_block_bwd[rows={B * T},C={Cc},I={I}]"
                        flops = 2.0 * B * T * (2 * Cc * I)
                    elif name == "osb_gemm_wgrad":
                        B, T, N, K, taps = args[5], args[6], args[7], args[8], args[9]
                        key = f"osb_gemm_wgrad[rows={B * T},N={N},K={K},taps={taps}]"
                        flops = 2.0 * B * T * N * K * taps
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    rc = fn(*args)
                    e1.record()
                    self.records.append((name, key, flops, e0, e1))
                    self.calls.append((fn, args))
                    return rc

                return wrapped

            setattr(lib, name, make())
        return self

    def __exit__(self, *exc):
        lib = load()
        for name, fn in self._saved.items():
            setattr(lib, name, fn)
        self._saved = {}

    def replay_amortized(self, repeat: int = 8, only_with_flops: bool = True) -> int:
        """Re-issues every recorded launch (same entry point, same arguments, same stream) `repeat` times back to back between
        ONE pair of events and keeps the bracket / repeat as that launch's device time.

        Why: a bracket around a single launch also holds the event records and the launch latency of an empty stream — about
        10 us on a 15-20 us kernel (tools/probe_gemm_graph.py: the same GEMM costs 17.9 us per launch inside a CUDA graph and
        30 us in a single bracket).  Call it right after the profiled step, before anything else allocates: the launches write
        into the finished step's (dead) activation / gradient buffers, which the caching allocator still owns; parameters and
        optimizer state are not among a contraction kernel's outputs.  Returns the number of launches replayed."""
        import torch

        torch.cuda.synchronize()
        torch.cuda._sleep(20_000_000)   # ~10 ms: the host queues the replays while the GPU is parked, brackets hold device time only
        ev = []
        for i, ((name, key, flops, _e0, _e1), (fn, args)) in enumerate(zip(self.records, self.calls)):
            if only_with_flops and not flops:
                continue
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(repeat):
                fn(*args)
            e1.record()
            ev.append((i, e0, e1))
        torch.cuda.synchronize()
        for i, e0, e1 in ev:
            self.amortized[i] = e0.elapsed_time(e1) / repeat
        return len(ev)

    def summary(self):
        """-> list of dicts sorted by total device time: key, launches, total_ms, avg_us, flops_per_launch; `total_ms` uses the
        amortized time of a launch where replay_amortized() produced one (`bracket_ms` keeps the single-bracket sum)."""
        import torch

        torch.cuda.synchronize()
        agg = {}
        for i, (name, key, flops, e0, e1) in enumerate(self.records):
            a = agg.setdefault(key, dict(key=key, name=name, launches=0, total_ms=0.0, bracket_ms=0.0, flops=0.0))
            a["launches"] += 1
            single = e0.elapsed_time(e1)
            a["bracket_ms"] += single
            a["total_ms"] += self.amortized.get(i, single)
            a["flops"] += flops
        out = sorted(agg.values(), key=lambda a: -a["total_ms"])
        for a in out:
            a["avg_us"] = 1e3 * a["total_ms"] / a["launches"]
            a["avg_us_single_bracket"] = 1e3 * a["bracket_ms"] / a["launches"]
            a["flops_per_launch"] = a["flops"] / a["launches"]
        return out
