"""CPU suite: the C-ABI library builds, loads and exports exactly what include/lxg.h declares;
without a GPU it fails loudly instead of falling back."""

import ctypes
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "lxg.h"


def header_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(lxg_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from lean_explore_b200 import _lib, build

    build.build()
    return _lib.load()


def test_header_and_ctypes_table_agree():
    from lean_explore_b200 import _lib

    assert header_functions() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol(lib):
    from lean_explore_b200 import _lib

    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True, check=True)
    exported = {line.split()[-1] for line in out.stdout.splitlines() if " T " in line}
    missing = [f for f in header_functions() if f not in exported]
    assert not missing, missing
    for name in header_functions():
        assert getattr(lib, name) is not None
    assert lib.lxg_abi_version() >= 1


def test_library_is_sm100a_with_tcgen05_and_tma():
    """The shipped cubin is sm_100a and its hot kernel uses tcgen05.mma / tcgen05.ld / TMA
    (SASS mnemonics from the B200 profiling guide)."""
    from lean_explore_b200 import _lib

    out = subprocess.run(["cuobjdump", "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    sass = out.stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic


def test_no_gpu_means_loud_failure_not_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lean_explore_b200 import _lib

    rc = lib.lxg_init(0)
    assert rc == -3  # LXG_ENODEVICE
    assert b"no CPU fallback" in lib.lxg_last_error()
    with pytest.raises(_lib.LxgError):
        _lib.init(0)
    from lean_explore_b200 import GpuIndexFlatIP

    with pytest.raises(_lib.LxgError):
        GpuIndexFlatIP(64)
    h = ctypes.c_void_p()
    assert lib.lxg_index_create(ctypes.byref(h), None, 0, 64, 1, 0) == -1  # not initialised -> EINVAL
    assert not h


def test_product_package_never_imports_the_oracle():
    for path in (ROOT / "lean_explore_b200").rglob("*.py"):
        text = path.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), path
    for path in (ROOT / "lean_explore_b200" / "csrc").iterdir():
        if path.suffix in (".cu", ".cuh", ".h"):
            assert "oracle/" not in path.read_text().replace("oracle/faiss_flat.py)", ""), path


def test_header_is_plain_c(tmp_path):
    """include/lxg.h is the whole boundary: it must compile as C99 on its own (no C++, no CUDA, no
    torch types) and every handle / struct it declares must be usable from C."""
    src = tmp_path / "use_lxg.c"
    src.write_text(
        '#include "lxg.h"\n'
        "int probe(void) {\n"
        "  lxg_index* ix = 0; lxg_encoder* enc = 0; lxg_decoder* dec = 0;\n"
        "  lxg_search_stats st; lxg_timing tm; lxg_bert_weights bw; lxg_qwen3_weights qw; lxg_qwen3_layer ql;\n"
        "  (void)ix; (void)enc; (void)dec; (void)st; (void)tm; (void)bw; (void)qw; (void)ql;\n"
        "  return LXG_OK + LXG_F16 + LXG_POOL_CLS + (int)sizeof(lxg_bert_layer);\n"
        "}\n")
    out = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only",
                          "-I", str(HEADER.parent), str(src)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
