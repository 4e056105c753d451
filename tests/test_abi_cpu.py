"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/osb200.h declares; the host layer refuses to run without CUDA tensors (no silent fallback)."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    from optispeech_b200 import _lib, build

    build.build()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    from optispeech_b200 import _lib

    names = _lib.exported_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/osb200.h but not exported by libosb200.so"


def test_header_is_plain_c():
    import subprocess
    import tempfile

    with tempfile.TemporaryDirectory() as d:
        src = Path(d) / "t.c"
        src.write_text('#include "osb200.h"\nint main(void){ osb_gemm_desc d; (void)d; return 0; }\n')
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), "-c", str(src), "-o", str(Path(d) / "t.o")],
                       check=True)


def test_gemm_desc_layout_matches_header(lib):
    """The ctypes mirror must have the same field order as the C struct."""
    from optispeech_b200 import _lib

    header = (ROOT / "include" / "osb200.h").read_text()
    body = header.split("typedef struct osb_gemm_desc {")[1].split("} osb_gemm_desc;")[0]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    c_fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.replace("*", " ").split()
        # "int32_t B, T, N, K" -> B T N K ; "const void* a" -> a
        first_type_tokens = 1 + (names[0] == "const")
        rest = " ".join(names[first_type_tokens:])
        c_fields += [n.strip() for n in rest.split(",")]
    assert c_fields == [f[0] for f in _lib.GemmDesc._fields_]


def test_version_and_errors(lib):
    assert lib.osb_version() >= 1
    assert b"shape" in lib.osb_strerror(-1)
    assert lib.osb_launch_count() == 0 or lib.osb_launch_count() > 0


def test_no_cpu_fallback():
    from optispeech_b200 import _lib, ops

    with pytest.raises(_lib.OsbError):
        ops.layernorm(torch.zeros(4, 256), torch.ones(256), torch.zeros(256), 1e-6)


def test_product_does_not_import_oracle():
    """Nothing under optispeech_b200/ may import the oracle (it is test infrastructure)."""
    for p in (ROOT / "optispeech_b200").rglob("*.py"):
        txt = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), p
