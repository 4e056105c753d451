"""Operand precision policy of the tensor-core path.

"fp16"   : operands rounded to fp16 once (2^-11 relative), fp32 accumulation - what the reference's own GPU
           default (`precision: 16-mixed`, configs/trainer/default.yaml:11) amounts to.  Used for training.
"fp16x3" : split precision (hi + lo fp16 parts, three tcgen05 passes, ~2^-21 relative).  Used for synthesis,
           where the north star asks for <= 1e-3 max-abs waveform error at full-scale amplitude.
Override with OSB_PRECISION=fp16|fp16x3 (applies to inference only; training always uses "fp16").
"""
from __future__ import annotations

import os

_INFERENCE_SPLIT = os.environ.get("OSB_PRECISION", "fp16x3").lower() != "fp16"


def set_inference_precision(name: str) -> None:
    global _INFERENCE_SPLIT
    if name not in ("fp16", "fp16x3"):
        raise ValueError(f"unknown precision {name!r}")
    _INFERENCE_SPLIT = name == "fp16x3"


def inference_precision() -> str:
    return "fp16x3" if _INFERENCE_SPLIT else "fp16"


def use_split(training: bool) -> bool:
    return _INFERENCE_SPLIT and not training
