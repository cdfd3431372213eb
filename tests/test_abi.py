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


def _plan(n, d, dtype, nq, k, sms=148):
    from lean_explore_b200 import _lib

    info = _lib.PlanInfo()
    lib = _lib.load()  # loading needs no GPU; lxg_debug_plan is pure host arithmetic
    assert lib.lxg_debug_plan(n, d, dtype, sms, nq, k, info) == 0
    return info


def test_pass1_plan_invariants_on_the_cpu():
    """The planning rules of DESIGN.md 4.1 (lxg_debug_plan runs make_plan without a device): query blocks,
    slices, the cross-list level's tracker depth / classes, list capacity - over the bench configs, the engine's
    request shape and a sweep of odd sizes."""
    F32, F16 = 0, 1
    cases = [(500_000, 384, F16, q, 50) for q in (1, 8, 64, 256, 512, 1024, 4096)]
    cases += [(2_000_000, 768, F16, 1024, 50), (16_000_000, 768, F16, 1024, 50), (50_000, 384, F32, 1000, 10),
              (400_000, 1024, F32, 1, 1000), (400_000, 1024, F32, 64, 1000), (400_000, 1024, F32, 1024, 50)]
    cases += [(n, d, F16, q, k) for n in (1, 100, 5_000, 77_777) for d in (64, 100, 1024) for q in (1, 129, 3000) for k in (1, 50, 2048)]
    for n, d, dt, q, k in cases:
        p = _plan(n, d, dt, q, k)
        assert p.kp >= k and p.kp % 32 == 0, (n, d, q, k)
        assert p.query_blocks == (q + 127) // 128
        assert 1 <= p.slices <= 148 and p.lists == 2 * p.slices
        grid_x = (p.query_blocks + 1) // 2 * 2 if p.pair else p.query_blocks
        assert p.pair == (1 if p.query_blocks >= 2 else 0)
        assert grid_x * p.slices <= 148 or p.slices == 1  # one co-resident CTA per SM
        assert p.tile_rows == 128
        if p.level_depth:
            assert p.level_depth in (1, 2, 4, 8) and p.level_classes in (1, 2, 4, 8) and p.level_classes <= p.level_depth
            ranks, weights = list(p.level_rank[: p.level_classes]), list(p.level_weight[: p.level_classes])
            assert ranks == sorted(set(ranks)) and ranks[-1] == p.level_depth and all(1 <= r <= 8 for r in ranks)
            assert sum(weights) == p.level_depth  # every row of a tracker is counted exactly once
            assert weights == [r - (ranks[i - 1] if i else 0) for i, r in enumerate(ranks)]
            assert p.lists * p.level_classes <= 1280  # words a level warp selects over
            assert p.lists * p.level_depth >= p.kp  # the union can hold kp entries
            assert p.list_capacity >= 1024 and p.merge_pool >= 4 * p.kp
        else:
            assert p.list_capacity >= p.kp + 2 * p.tile_rows  # per-list compaction: room for a tile between two compactions
    # the shapes DESIGN.md quotes
    p = _plan(500_000, 384, F16, 1024, 50)
    assert (p.kp, p.slices, p.lists, p.pair, p.level_depth, p.level_classes) == (64, 18, 36, 1, 4, 4)
    p = _plan(500_000, 384, F16, 1, 1000)
    assert (p.kp, p.level_depth, p.level_classes) == (1280, 8, 4) and list(p.level_rank[:4]) == [1, 2, 4, 8] and list(p.level_weight[:4]) == [1, 1, 2, 4]
    p = _plan(500_000, 384, F16, 4096, 50)
    assert (p.lists, p.level_depth, p.level_classes) == (8, 8, 8)  # union == kp: the cheap level is the level
    assert _plan(400_000, 1024, F32, 1, 1000).kp == 1408
