"""In-tree build of libosb200.so (the C-ABI CUDA library) with nvcc for sm_100a.

The shared object is written next to this file so that it travels with the repo snapshot
to the GPU box (it is git-ignored, not gpurun-ignored).  There is deliberately no JIT and no
fallback: if the library is missing the product path raises.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
INCLUDE = PKG_DIR.parent / "include"
LIB_PATH = PKG_DIR / "libosb200.so"
STAMP_PATH = PKG_DIR / "libosb200.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
]


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list(INCLUDE.glob("*.h"))):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libosb200.so cannot be built")
    return nvcc


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every csrc/*.cu into libosb200.so (one object per file, built in parallel)."""
    fp = _fingerprint()
    if not force and LIB_PATH.exists() and STAMP_PATH.exists() and STAMP_PATH.read_text().strip() == fp:
        return LIB_PATH
    nvcc = find_nvcc()
    obj_dir = PKG_DIR / "build"
    obj_dir.mkdir(exist_ok=True)
    procs = []
    objs = []
    for src in _sources():
        obj = obj_dir / (src.stem + ".o")
        objs.append(str(obj))
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(INCLUDE), "-I", str(CSRC), "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, proc in procs:
        out, _ = proc.communicate()
        if verbose or proc.returncode != 0:
            print(f"--- {src.name}\n{out}")
        failed |= proc.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed; see output above")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH), *objs]
    res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        print(res.stdout)
        raise RuntimeError("link of libosb200.so failed")
    STAMP_PATH.write_text(fp)
    return LIB_PATH


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
