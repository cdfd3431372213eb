#!/usr/bin/env python
"""Generates tests/golden/flat_ip_golden.npz.

The reference's arithmetic for this path lives in the faiss-cpu wheel, which cannot be installed
here (no network) and is not under /root/reference, so these vectors cannot come from the
reference itself: they are the exact-arithmetic (fp64) ranking computed by
oracle/faiss_flat.py, frozen so that later changes to the oracle or to numpy are detected.
Inputs are regenerated from seeds (tests/conftest.py); a SHA-256 of the input bytes is stored
with every case so a drifting generator is detected rather than silently re-baselined.
The one case that does come from the reference is its known-answer test
(/root/reference/tests/extract/index_test.py:185-205), stored as `kat_*`.

Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from conftest import make_corpus, make_queries  # noqa: E402
from oracle import faiss_flat as ff  # noqa: E402

CASES = [
    # name, n, d, nq, k, corpus dtype, normalize
    ("n4096_d384_q32_k10_f16", 4096, 384, 32, 10, "float16", True),
    ("n4096_d384_q32_k50_f16", 4096, 384, 32, 50, "float16", True),
    ("n4096_d768_q32_k1_f16", 4096, 768, 32, 1, "float16", True),
    ("n4096_d768_q32_k50_f32", 4096, 768, 32, 50, "float32", True),
    ("n1001_d100_q5_k7_f32_raw", 1001, 100, 5, 7, "float32", False),
    ("n7_d64_q3_k12_f16_pad", 7, 64, 3, 12, "float16", True),
]


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    out = {}
    for name, n, d, nq, k, dtype, normalize in CASES:
        c = make_corpus(n, d, dtype=np.dtype(dtype).type)
        x = make_queries(nq, d)
        xn = x.copy()
        if normalize:
            ff.normalize_L2(xn)
        D, I = ff.flat_ip_search_f64(c, xn, k)
        out[name + "/I"] = I.astype(np.int32)
        out[name + "/D"] = D
        out[name + "/sha"] = np.array(digest(c, x))
        out[name + "/cfg"] = np.array([n, d, nq, k, int(dtype == "float32"), int(normalize)])
    # reference KAT: one-hot row 0 is its own nearest neighbour
    out["kat/I"] = np.array([[0]], dtype=np.int32)
    out["kat/D"] = np.array([[1.0]])
    np.savez_compressed(Path(__file__).with_name("flat_ip_golden.npz"), **out)
    print("wrote", len(CASES), "cases")


if __name__ == "__main__":
    main()
