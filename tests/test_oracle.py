"""CPU suite: the oracle (oracle/faiss_flat.py, oracle/flat_ip.c) against the committed golden
vectors, the reference's known-answer test and FAISS' documented edge-case behaviour."""

import numpy as np
import pytest

from conftest import make_corpus, make_queries
from golden_cases import case_names, load_case
from oracle import c_oracle
from oracle import faiss_flat as ff


@pytest.mark.parametrize("name", case_names())
def test_oracle_matches_golden(name):
    g = load_case(name)
    xn = g["x"].copy()
    if g["normalize"]:
        ff.normalize_L2(xn)
    D, I = ff.flat_ip_search_f64(g["corpus"], xn, g["k"])
    assert np.array_equal(I, g["I"])
    live = I >= 0
    assert np.allclose(D[live], g["D"][live], rtol=0, atol=1e-12)
    # the fp32 (sgemm) flavour agrees except at fp64-certified near ties
    D32, I32 = ff.flat_ip_search(g["corpus"], xn, g["k"])
    differ = I32 != I
    scale = float(np.linalg.norm(xn, axis=1).max() * np.linalg.norm(g["corpus"].astype(np.float32), axis=1).max())
    assert ff.ambiguous_positions(g["D"], tol=4e-6 * scale)[differ].all()
    assert np.abs(D32[live] - g["D"][live]).max() < 1e-3 * max(1.0, scale)


def test_reference_kat_one_hot_row():
    """reference tests/extract/index_test.py:185-205: a one-hot row is its own nearest neighbour."""
    emb = np.random.default_rng(123).random((300, 768), dtype=np.float32)
    emb[0] = 0.0
    emb[0, 0] = 1.0
    q = np.zeros((1, 768), dtype=np.float32)
    q[0, 0] = 1.0
    ix = ff.IndexFlatIP(768)
    ix.add(emb)
    assert ix.ntotal == 300 and ix.d == 768
    D, I = ix.search(q, 1)
    assert I[0][0] == 0 and D[0][0] == 1.0
    Dc, Ic = c_oracle.knn_inner_product_seq(q, emb, 1)
    assert Ic[0][0] == 0 and Dc[0][0] == 1.0


@pytest.mark.parametrize("nq", [5, 40])  # FAISS switches from per-pair SIMD to sgemm at nq = 20
def test_python_and_c_oracle_agree_on_tie_free_data(nq):
    c = make_corpus(3000, 96, dtype=np.float32)
    x = make_queries(nq, 96)
    ff.normalize_L2(x)
    D64, I64 = ff.flat_ip_search_f64(c, x, 10)
    fn = c_oracle.knn_inner_product_seq if nq < 20 else c_oracle.knn_inner_product_blas
    Dc, Ic = fn(x, c, 10)
    differ = Ic != I64
    assert ff.ambiguous_positions(D64, tol=4e-6)[differ].all()
    assert differ.mean() < 0.01
    assert np.abs(Dc - D64).max() < 1e-5


def test_k_larger_than_n_pads_minus_one_and_lowest_float():
    c = make_corpus(7, 64)
    x = make_queries(3, 64)
    D, I = ff.flat_ip_search(c, x, 12)
    assert (I[:, 7:] == -1).all() and (D[:, 7:] == ff.NEG_FLT_MAX).all()
    assert (np.sort(I[:, :7], axis=1) == np.arange(7)).all()
    Dc, Ic = c_oracle.knn_inner_product_seq(x, c.astype(np.float32), 12)
    assert np.array_equal(Ic, I)
    assert (Dc[:, 7:] == ff.NEG_FLT_MAX).all()


def test_results_sorted_descending_and_ties_by_row_id():
    c = make_corpus(500, 32, dtype=np.float32)
    c[100:110] = c[3]
    x = c[[3]].copy()
    D, I = ff.flat_ip_search_f64(c, x, 12)
    assert np.all(np.diff(D[0]) <= 0)
    assert list(I[0][:11]) == [3] + list(range(100, 110))


def test_normalize_l2_semantics():
    x = make_queries(9, 384) * 5
    x[2] = 0.0
    want = x.copy()
    ff.normalize_L2(want)
    assert np.all(want[2] == 0.0)  # zero rows untouched
    norms = np.linalg.norm(want[[0, 1, 3, 4, 5, 6, 7, 8]].astype(np.float64), axis=1)
    assert np.abs(norms - 1).max() < 1e-6
    got = x.copy()
    c_oracle.renorm_l2(got)  # fp32 SIMD-order accumulation, as FAISS
    assert np.abs(got - want).max() < 2e-7
    with pytest.raises(TypeError):
        ff.normalize_L2(x.astype(np.float64))


def test_retrieve_semantic_candidates_glue():
    """engine.py:238-258: -1 and out-of-range labels skipped, max similarity per declaration."""

    class FakeIndex:
        def search(self, q, k):
            return (np.array([[0.9, 0.8, 0.7, 0.6, -3.4e38]], dtype=np.float32),
                    np.array([[2, 0, 5, 1, -1]], dtype=np.int64))

    id_map = [10, 11, 10]  # rows 0 and 2 belong to the same declaration; label 5 is out of range
    got = ff.retrieve_semantic_candidates(FakeIndex(), id_map, [0.1] * 8, 5)
    assert got == {10: pytest.approx(0.9), 11: pytest.approx(0.6)}


def test_empty_corpus():
    D, I = ff.flat_ip_search(np.zeros((0, 16), dtype=np.float32), make_queries(2, 16), 3)
    assert (I == -1).all() and (D == ff.NEG_FLT_MAX).all()
