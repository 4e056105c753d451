"""Recipe: stage the reference's Python package under oracle/_ref/ (git-ignored) so that it travels to the GPU box.

    python -m oracle.make_ref            # build container only: needs /root/reference (read-only, never modified)

TEST INFRASTRUCTURE.  The reference is pure Python (no build step); what is copied is the sub-tree the hot path imports:
`optispeech/{__init__,values}.py`, `optispeech/model/**`, `optispeech/utils/**`, `optispeech/text/**` (~0.7 MB; the 21 MB
`vendor/` tree, `dataset/`, `onnx/`, `tools/` are not on the path).  Nothing under oracle/_ref/ is ever committed
(.gitignore) or imported by the product package; bench.py's `--impl reference` / `--impl torch-gpu` legs and the golden
generators import it through oracle/ref_harness.py.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/optispeech"
DST = os.path.join(HERE, "_ref", "optispeech")
KEEP = ("__init__.py", "values.py", "model", "utils", "text")


def make_ref(verbose: bool = True) -> str | None:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"make_ref: {SRC} not present (GPU box?) - keeping whatever oracle/_ref already holds")
        return DST if os.path.isdir(DST) else None
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    for name in KEEP:
        s, d = os.path.join(SRC, name), os.path.join(DST, name)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.onnx", "*.pt", "*.ckpt"))
        elif os.path.isfile(s):
            shutil.copy2(s, d)
    with open(os.path.join(HERE, "_ref", "README"), "w") as f:
        f.write("Unmodified copy of /root/reference/optispeech/{__init__.py,values.py,model,utils,text} made by oracle/make_ref.py.\n"
                "Git-ignored; used only as the checker / reference arm (never shipped, never imported by optispeech_b200).\n")
    if verbose:
        n = sum(len(fs) for _, _, fs in os.walk(DST))
        print(f"make_ref: staged {n} files under {DST}")
    return DST


if __name__ == "__main__":
    sys.exit(0 if make_ref() else 1)
