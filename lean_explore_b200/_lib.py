"""ctypes binding of ``liblxg.so`` (the C ABI declared in ``include/lxg.h``).

There is deliberately no fallback: if the shared library is missing, or no sm_100 GPU is
visible, every entry point of this package raises.
"""

from __future__ import annotations

import ctypes
import threading
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p
from pathlib import Path

import os

# LXG_LIB_PATH selects an A/B build variant of the same ABI (lean_explore_b200.build.build(defines=...))
LIB_PATH = Path(os.environ.get("LXG_LIB_PATH") or Path(__file__).resolve().parent / "liblxg.so")

LXG_F32, LXG_F16 = 0, 1
LXG_POOL_MEAN, LXG_POOL_CLS = 0, 1


class LxgError(RuntimeError):
    """A C-ABI call returned a negative status (message from ``lxg_last_error``)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"lxg error {code}: {message}")
        self.code = code


class SearchStats(ctypes.Structure):
    _fields_ = [
        ("kernel_launches", c_int32),
        ("slices", c_int32),
        ("query_blocks", c_int32),
        ("kp", c_int32),
        ("uncertified", c_int32),
        ("tile_rows", c_int32),
    ]


class PlanInfo(ctypes.Structure):
    """lxg_plan_info: the pass-1 layout lxg_debug_plan reports (include/lxg.h)."""
    _fields_ = [
        ("kp", c_int32), ("query_blocks", c_int32), ("slices", c_int32), ("lists", c_int32), ("tile_rows", c_int32),
        ("pair", c_int32), ("level_depth", c_int32), ("level_classes", c_int32), ("level_rank", c_int32 * 8),
        ("level_weight", c_int32 * 8), ("list_capacity", c_int32), ("merge_pool", c_int32),
    ]


class Timing(ctypes.Structure):
    _fields_ = [("calls", c_int32), ("scan_ms", c_float), ("merge_ms", c_float), ("exact_ms", c_float),
                ("prep_ms", c_float)]


class BertLayer(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in (
        "wqkv", "bqkv", "wo", "bo", "ln1_g", "ln1_b", "w1", "b1", "w2", "b2", "ln2_g", "ln2_b")]


class BertWeights(ctypes.Structure):
    _fields_ = [
        ("hidden", c_int32), ("layers", c_int32), ("heads", c_int32), ("ffn", c_int32),
        ("vocab", c_int32), ("max_pos", c_int32), ("ln_eps", c_float),
        ("word_emb", c_void_p), ("pos_emb", c_void_p), ("type_emb", c_void_p),
        ("emb_ln_g", c_void_p), ("emb_ln_b", c_void_p),
        ("layer", POINTER(BertLayer)),
    ]


class Qwen3Layer(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in ("ln1", "wqkv", "q_norm", "k_norm", "wo", "ln2", "wgu", "wdown")]


class Qwen3Weights(ctypes.Structure):
    _fields_ = [
        ("hidden", c_int32), ("layers", c_int32), ("heads", c_int32), ("kv_heads", c_int32),
        ("head_dim", c_int32), ("ffn", c_int32), ("vocab", c_int32), ("rms_eps", c_float),
        ("tok_emb", c_void_p), ("lm_head", c_void_p), ("final_norm", c_void_p), ("inv_freq", c_void_p),
        ("layer", POINTER(Qwen3Layer)),
    ]


# name -> (restype, argtypes); mirrors include/lxg.h one to one (tests/test_abi.py checks it)
SIGNATURES = {
    "lxg_init": (c_int, [c_int]),
    "lxg_last_error": (c_char_p, []),
    "lxg_abi_version": (c_int, []),
    "lxg_index_create": (c_int, [POINTER(c_void_p), c_void_p, c_int64, c_int32, c_int, c_int64]),
    "lxg_index_destroy": (c_int, [c_void_p]),
    "lxg_index_ntotal": (c_int64, [c_void_p]),
    "lxg_index_d": (c_int32, [c_void_p]),
    "lxg_search": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int, c_void_p, c_void_p, c_void_p]),
    "lxg_search_ex": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int, c_void_p, c_void_p,
                              c_void_p, c_void_p]),
    "lxg_normalize_l2": (c_int, [c_void_p, c_int32, c_int32, c_void_p]),
    "lxg_merge_topk": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                               c_void_p]),
    "lxg_merge_topk_packed": (c_int, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "lxg_index_sync": (c_int, [c_void_p, POINTER(c_int32)]),
    "lxg_index_last_stats": (c_int, [c_void_p, POINTER(SearchStats)]),
    "lxg_index_set_timing": (c_int, [c_void_p, c_int]),
    "lxg_index_get_timing": (c_int, [c_void_p, POINTER(Timing)]),
    "lxg_debug_scores": (c_int, [c_void_p, c_void_p, c_int32, c_int, c_void_p, c_void_p,
                                 POINTER(c_float), c_void_p]),
    "lxg_debug_config": (c_int, [c_int, c_int, c_int]),
    "lxg_debug_plan": (c_int, [c_int64, c_int32, c_int, c_int32, c_int32, c_int32, POINTER(PlanInfo)]),
    "lxg_encoder_create": (c_int, [POINTER(c_void_p), POINTER(BertWeights)]),
    "lxg_encoder_destroy": (c_int, [c_void_p]),
    "lxg_encoder_last_launches": (c_int, [c_void_p]),
    "lxg_encoder_set_fused": (c_int, [c_void_p, c_int]),
    "lxg_encoder_read_trace": (c_int, [c_void_p, c_void_p, c_int32, POINTER(c_int32), POINTER(c_int32)]),
    "lxg_encode": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int, c_void_p, c_void_p]),
    "lxg_decoder_create": (c_int, [POINTER(c_void_p), POINTER(Qwen3Weights)]),
    "lxg_decoder_destroy": (c_int, [c_void_p]),
    "lxg_decoder_last_launches": (c_int, [c_void_p]),
    "lxg_decoder_last_tokens": (c_int, [c_void_p]),
    "lxg_decoder_embed": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "lxg_decoder_rerank": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                   c_void_p]),
}

_lock = threading.Lock()
_lib = None
_inited_devices: set[int] = set()


def load() -> ctypes.CDLL:
    """dlopen liblxg.so and declare every prototype.  Raises if the library is not built."""
    global _lib
    with _lock:
        if _lib is None:
            if not LIB_PATH.exists():
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with `python -m lean_explore_b200.build` "
                    "(nvcc, sm_100a). There is no CPU fallback for the lxg kernels."
                )
            lib = ctypes.CDLL(str(LIB_PATH))
            for name, (restype, argtypes) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = restype
                fn.argtypes = argtypes
            _lib = lib
        return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().lxg_last_error()
        raise LxgError(rc, msg.decode() if msg else "")


def init(device: int = 0) -> ctypes.CDLL:
    """Load the library and bring up `device` (must be an sm_100 GPU)."""
    lib = load()
    if device not in _inited_devices:
        # lxg_init makes `device` current on the calling thread; every later entry point switches to
        # its handle's device by itself, so restore what the caller (torch) had selected
        import torch

        prev = torch.cuda.current_device() if torch.cuda.is_available() else None
        check(lib.lxg_init(device))
        if prev is not None and prev != device:
            torch.cuda.set_device(prev)
        with _lock:
            _inited_devices.add(device)
    return lib
