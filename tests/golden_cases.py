"""Loader of tests/golden/flat_ip_golden.npz shared by the CPU (oracle) and GPU (parity) suites."""
import hashlib
from pathlib import Path

import numpy as np

from conftest import make_corpus, make_queries

GOLDEN = Path(__file__).resolve().parent / "golden" / "flat_ip_golden.npz"


def case_names():
    z = np.load(GOLDEN)
    return sorted({k.split("/")[0] for k in z.files if not k.startswith("kat")})


def load_case(name):
    """-> dict(corpus, x, k, normalize, I (int64), D (float64)); checks the input digest."""
    z = np.load(GOLDEN)
    n, d, nq, k, is_f32, normalize = (int(v) for v in z[name + "/cfg"])
    c = make_corpus(n, d, dtype=np.float32 if is_f32 else np.float16)
    x = make_queries(nq, d)
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(c).tobytes())
    h.update(np.ascontiguousarray(x).tobytes())
    assert h.hexdigest() == str(z[name + "/sha"]), "synthetic-input generator drifted; golden vectors no longer apply"
    return dict(corpus=c, x=x, k=k, normalize=bool(normalize), I=z[name + "/I"].astype(np.int64), D=z[name + "/D"])
