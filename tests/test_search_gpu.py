"""GPU parity tests of the flat inner-product search (through the C ABI) against the CPU
oracle.  Shapes follow BASELINE.json configs at sizes the oracle finishes in seconds; edge
cases follow SURVEY.md section 8(c) and the reference's one KAT."""

import numpy as np
import pytest
import torch

from conftest import make_corpus, make_queries
from oracle import faiss_flat as ff

pytestmark = pytest.mark.gpu


def _index(corpus_np):
    from lean_explore_b200 import GpuIndexFlatIP

    return GpuIndexFlatIP.from_tensor(torch.from_numpy(corpus_np).cuda())


def _check(corpus_np, x, k, normalize=True):
    """ids must equal the exact-arithmetic ranking; scores within 1e-3 (north star) of fp32 FAISS
    semantics; disagreement with the fp32 oracle only at fp64-certified near-ties."""
    ix = _index(corpus_np)
    D, I = ix.search(x, k, normalize=normalize)
    xn = x.copy()
    if normalize:
        ff.normalize_L2(xn)
    D64, I64 = ff.flat_ip_search_f64(corpus_np, xn, k)
    D32, I32 = ff.flat_ip_search(corpus_np, xn, k)
    assert I.dtype == np.int64 and D.dtype == np.float32
    assert np.array_equal(I, I64), f"{(I != I64).sum()} ids differ from the exact ranking"
    live = I64 >= 0
    mag = max(1.0, float(np.abs(D64[live]).max())) if live.any() else 1.0
    assert np.abs(D[live] - D64[live]).max() < 1e-3 * mag  # tolerance stated by the north star
    assert np.all(D[~live] == ff.NEG_FLT_MAX)
    differ = I != I32
    if differ.any():
        # fp32 sgemm noise scales with ||q|| * ||c||: only there may two fp32 FAISS builds disagree
        scale = float(np.linalg.norm(xn, axis=1).max() * np.linalg.norm(corpus_np.astype(np.float32), axis=1).max())
        assert ff.ambiguous_positions(D64, tol=4e-6 * scale)[differ].all(), "differs from fp32 oracle away from a near-tie"
        assert differ.sum() <= max(4, 1e-3 * differ.size)  # a swapped near-tie pair costs 2 positions
    return ix


@pytest.mark.parametrize("n,d,nq", [(300, 64, 5), (1000, 384, 7), (1000, 768, 130), (1001, 128, 3),
                                     (1000, 1024, 130), (700, 1000, 5), (900, 840, 260)])
def test_tensor_core_scores_match_reference(n, d, nq):
    """Raw pass-1 scores (TMA -> swizzled smem -> tcgen05.mma with A in TMEM -> tcgen05.ld)
    against an fp64 product of the same fp16-rounded operands."""
    corpus = make_corpus(n, d)
    x = make_queries(nq, d)
    ix = _index(corpus)
    xt = torch.from_numpy(x).cuda()
    got = ix.debug_scores(xt, normalize=True).cpu().numpy()
    xn = x.copy()
    ff.normalize_L2(xn)
    ref = xn.astype(np.float64) @ corpus.astype(np.float64).T
    err = np.abs(got - ref).max()
    assert err < 2e-3, f"max abs err {err}"
    # and with the query rounded the way the kernel rounds it the match is to fp32 accumulate noise
    amax = np.abs(xn).max(axis=1, keepdims=True)
    scale = 2.0 ** (-np.floor(np.log2(amax)))
    xh = (xn * scale).astype(np.float16).astype(np.float64) / scale
    ref16 = xh @ corpus.astype(np.float64).T
    assert np.abs(got - ref16).max() < 2e-5


def test_reference_kat_one_hot_row_is_its_own_neighbour():
    """The reference's only known-answer test (tests/extract/index_test.py:185-205)."""
    rng = np.random.default_rng(123)
    emb = rng.random((300, 768), dtype=np.float32)
    emb[0] = 0.0
    emb[0, 0] = 1.0
    q = np.zeros((1, 768), dtype=np.float32)
    q[0, 0] = 1.0
    ix = _index(emb)
    assert ix.ntotal == 300 and ix.d == 768
    D, I = ix.search(q, 1)
    assert I[0][0] == 0 and abs(D[0][0] - 1.0) < 1e-6


@pytest.mark.parametrize("dtype", [np.float16, np.float32])
@pytest.mark.parametrize("n,d,nq,k", [(4096, 384, 32, 10), (4096, 768, 32, 50), (20000, 384, 200, 50),
                                       (4096, 1024, 32, 50), (20000, 1024, 300, 50)])
def test_topk_matches_oracle(n, d, nq, k, dtype):
    _check(make_corpus(n, d, dtype=dtype), make_queries(nq, d), k)


def test_config1_shape_fp32_corpus_top10():
    """BASELINE.json config 1: 1k-query batch, 50k x 384 fp32 corpus, top-10, ids bit-exact."""
    _check(make_corpus(50000, 384, dtype=np.float32), make_queries(1000, 384), 10)


def test_unnormalised_search_and_large_k():
    corpus = make_corpus(3000, 256)
    x = make_queries(9, 256) * 37.5
    _check(corpus, x, 200, normalize=False)
    _check(corpus, x, 1000, normalize=True)


def test_k_larger_than_n_pads_like_faiss():
    corpus = make_corpus(7, 64)
    x = make_queries(3, 64)
    ix = _check(corpus, x, 12)
    D, I = ix.search(x, 12, normalize=True)
    assert (I[:, 7:] == -1).all() and (D[:, 7:] == ff.NEG_FLT_MAX).all()
    assert (np.sort(I[:, :7], axis=1) == np.arange(7)).all()


def test_ragged_shapes():
    for n, d, nq in [(1, 8, 1), (129, 72, 2), (257, 100, 129), (5000, 760, 3), (5000, 1016, 3), (3000, 900, 140)]:
        _check(make_corpus(n, d), make_queries(nq, d), 5)
    _check(make_corpus(515, 100, dtype=np.float32), make_queries(4, 100), 5)


def test_empty_index_and_zero_queries():
    from lean_explore_b200 import GpuIndexFlatIP

    ix = GpuIndexFlatIP(64)
    D, I = ix.search(make_queries(2, 64), 3)
    assert (I == -1).all() and (D == ff.NEG_FLT_MAX).all()
    ix.add(make_corpus(10, 64))
    assert ix.ntotal == 10
    D, I = ix.search(np.zeros((0, 64), dtype=np.float32), 3)
    assert D.shape == (0, 3) and I.shape == (0, 3)
    # the documented limit: a clear ValueError, not a status code from deep inside
    with pytest.raises(ValueError, match="at most 2048"):
        ix.search(make_queries(1, 64), 2049)
    with pytest.raises(ValueError, match="at most 2048"):
        ix.search_torch(torch.zeros((1, 64), device="cuda"), 4096)


def test_zero_norm_query_row():
    """normalize_L2 leaves an all-zero row untouched; every score is 0 and ties go to the lowest ids."""
    corpus = make_corpus(2000, 128)
    x = make_queries(3, 128)
    x[1] = 0.0
    ix = _index(corpus)
    D, I = ix.search(x, 10, normalize=True)
    assert np.array_equal(I[1], np.arange(10)) and (D[1] == 0).all()
    xn = x.copy()
    ff.normalize_L2(xn)
    _, I64 = ff.flat_ip_search_f64(corpus, xn, 10)
    assert np.array_equal(I, I64)


def test_duplicate_rows_tie_break_by_row_id():
    """Exact ties (duplicated rows) are ordered by ascending row id, also across slices."""
    corpus = make_corpus(30000, 128)
    corpus[20000:20100] = corpus[17]          # 100 copies of row 17 far away (more than kp)
    corpus[5:9] = corpus[29999]               # and a small group at the front
    x = corpus[[17, 29999, 3]].astype(np.float32)
    ix = _check(corpus, x, 20)
    D, I = ix.search(x, 20, normalize=True)
    assert I[0][0] == 17 and np.array_equal(I[0][1:20], np.arange(20000, 20019))
    assert np.array_equal(I[1][:5], [5, 6, 7, 8, 29999])
    assert ix.last_stats()["uncertified"] >= 1  # the tie group straddles rank k: exact path used


def test_normalize_l2_matches_oracle():
    from lean_explore_b200 import normalize_L2

    x = make_queries(33, 384) * 3.0
    x[4] = 0.0
    want = x.copy()
    ff.normalize_L2(want)
    got = x.copy()
    normalize_L2(got)
    assert np.array_equal(got, want)


def test_device_resident_search_and_stats():
    corpus = make_corpus(8192, 384)
    x = make_queries(256, 384)
    ix = _index(corpus)
    xt = torch.from_numpy(x).cuda()
    D, I = ix.search_torch(xt, 50, normalize=True)
    torch.cuda.synchronize()
    Dh, Ih = ix.search(x, 50, normalize=True)
    assert np.array_equal(I.cpu().numpy(), Ih) and np.array_equal(D.cpu().numpy(), Dh)
    st = ix.last_stats()
    assert st["kernel_launches"] == 5 and st["query_blocks"] == 2 and st["tile_rows"] == 128


def test_merge_topk_shards():
    """Row-sharded search (two half indexes + lxg_merge_topk) equals the single-index search."""
    from lean_explore_b200 import GpuIndexFlatIP, _lib

    corpus = make_corpus(6000, 256)
    x = make_queries(40, 256)
    k = 25
    full = _index(corpus)
    Df, If = full.search(x, k, normalize=True)
    ct = torch.from_numpy(corpus).cuda()
    halves = [GpuIndexFlatIP.from_tensor(ct[:2500].contiguous(), 0),
              GpuIndexFlatIP.from_tensor(ct[2500:].contiguous(), 2500)]
    xt = torch.from_numpy(x).cuda()
    parts = [h.search_torch(xt, k, normalize=True, want_f64=True) for h in halves]
    Dg = torch.stack([p[2] for p in parts]).contiguous()
    Ig = torch.stack([p[1] for p in parts]).contiguous()
    D = torch.empty((40, k), dtype=torch.float32, device="cuda")
    I = torch.empty((40, k), dtype=torch.int64, device="cuda")
    lib = _lib.init(0)
    _lib.check(lib.lxg_merge_topk(Dg.data_ptr(), Ig.data_ptr(), 40, k, 2, D.data_ptr(), I.data_ptr(), None))
    torch.cuda.synchronize()
    assert np.array_equal(I.cpu().numpy(), If) and np.array_equal(D.cpu().numpy(), Df)


@pytest.fixture
def scan_modes():
    """Toggle pass-1 code paths through the lxg_debug_config test hook; always restored."""
    from lean_explore_b200 import _lib

    lib = _lib.init(0)
    yield lambda **kw: _lib.check(lib.lxg_debug_config(kw.get("no_level", -1), kw.get("force_single", -1), -1))
    _lib.check(lib.lxg_debug_config(0, 0, 0))


def test_every_scan_mode_returns_the_same_exact_result(scan_modes):
    """Cross-list level vs compaction-only thresholds, CTA pairs vs single CTAs: four code paths of
    pass 1, one answer (ids exact against the oracle in each)."""
    corpus = make_corpus(30000, 384)
    x = make_queries(300, 384)
    want = None
    for no_level in (0, 1):
        for single in (0, 1):
            scan_modes(no_level=no_level, force_single=single)
            ix = _check(corpus, x, 50)
            D, I = ix.search(x, 50, normalize=True)
            if want is None:
                want = (D, I)
            assert np.array_equal(I, want[1]) and np.array_equal(D, want[0])


@pytest.mark.parametrize("n,d,nq,k,dtype", [(60000, 384, 1, 1000, np.float16), (40000, 1024, 1, 1000, np.float32),
                                           (30000, 256, 5, 400, np.float16), (3000, 128, 8, 2048, np.float32),
                                           (50000, 64, 2, 300, np.float16)])
def test_few_queries_large_k_three_stage_merge(n, d, nq, k, dtype):
    """The engine's own request (one query, faiss_k = 1000, engine.py:538) and its neighbours: up to 8
    queries with k' >= 256 run the merge as select -> re-score on every SM -> rank.  Exact ids
    against the oracle, and bit-identical results (scores included) to the one-kernel merge."""
    from lean_explore_b200 import _lib

    lib = _lib.init(0)
    corpus = make_corpus(n, d, dtype=dtype)
    x = make_queries(nq, d)
    x[nq // 2] *= 37.0  # un-normalised queries: the fused normalise is part of the call
    ix = _check(corpus, x, k)
    D3, I3 = ix.search(x, k, normalize=True)
    assert ix.last_stats()["kernel_launches"] == 7  # prep, scan, 3 merge stages, 2 exact-path kernels
    try:
        _lib.check(lib.lxg_debug_config(-1, -1, 16))
        D1, I1 = ix.search(x, k, normalize=True)
        assert ix.last_stats()["kernel_launches"] == 5
    finally:
        _lib.check(lib.lxg_debug_config(-1, -1, 17))
    assert np.array_equal(I1, I3) and np.array_equal(D1, D3)


def test_three_stage_merge_zero_query_duplicates_and_sorted_corpus():
    """Paths of stage 1 that finish a query by themselves (all-zero query; more near-ties than the pool
    holds -> exact path) and a row order that overflows the pool (radix tightening of the level)."""
    rng = np.random.default_rng(3)
    corpus = make_corpus(40000, 128)
    # rows sorted by their score against query 0: every list keeps appending, the pool overflows
    x = make_queries(3, 128)
    order = np.argsort(corpus.astype(np.float32) @ x[0])
    corpus = np.ascontiguousarray(corpus[order])
    x[1] = 0.0
    _check(corpus, x, 600)
    dup = make_corpus(20000, 96)
    dup[5000:9000] = dup[4999]  # 4001 identical rows: ties far beyond k
    x2 = np.stack([dup[4999].astype(np.float32) + 0.01 * rng.standard_normal(96).astype(np.float32), make_queries(1, 96)[0]])
    _check(dup, x2, 500)


@pytest.mark.parametrize("d", [64, 768, 1024])
def test_adversarial_row_order_overflows_the_lists(d):
    """Rows sorted by ascending score for the probe query: every tile beats everything seen before,
    thresholds always lag, the candidate lists overflow and are compacted exactly."""
    n = 60000
    corpus = make_corpus(n, d).astype(np.float32)
    q = make_queries(2, d)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    order = np.argsort(corpus @ q[0], kind="stable")
    corpus = np.ascontiguousarray(corpus[order]).astype(np.float16)
    _check(corpus, q, 10)     # 1 query block: many short lists, pass 2 has to tighten the level
    _check(corpus, q, 100)
    many = np.concatenate([q, make_queries(1198, d)])  # 10 query blocks: few long lists overflow
    _check(corpus, many, 10)


def test_shipped_model_dimension_d1024_fp32_corpus():
    """d = 1024 is what the shipped index holds (Qwen3-Embedding-0.6B): query chunks 8..15 of the
    block live in shared memory (SS MMAs), fp32 corpus as the reference stores it, faiss_k = 1000."""
    corpus = make_corpus(30000, 1024, dtype=np.float32)
    _check(corpus, make_queries(1, 1024), 1000)
    _check(corpus, make_queries(257, 1024), 10)
    # fp32 corpus at d = 1024: the scan copy is rounded too, so the certificate's eps doubles while
    # the scores concentrate (sigma = 1/32); the candidate margin k' - k grows with rel_err sqrt(d)
    # so that a batch does not fall back to the exhaustive exact path
    big = make_corpus(100000, 1024, dtype=np.float32)
    ix = _check(big, make_queries(256, 1024), 50)
    st = ix.last_stats()
    assert st["kp"] >= 96 and 0 <= st["uncertified"] <= 2, st
    from lean_explore_b200 import _lib

    with pytest.raises(_lib.LxgError):
        _index(make_corpus(10, 1032))


def test_large_k_with_many_query_blocks():
    """k' too large for the cross-list level (lists * 8 < k'): compaction thresholds, pair mode."""
    corpus = make_corpus(20000, 256)
    x = make_queries(300, 256)
    _check(corpus, x, 200)
    _check(corpus[:5000], x[:130], 1000)


def test_near_duplicate_heavy_corpus_goes_through_the_exact_path():
    """Many exact duplicates around rank k for several queries: uncertified -> exact collectors."""
    corpus = make_corpus(20000, 128)
    for j in range(8):
        corpus[1000 * j + 7 : 1000 * j + 47] = corpus[j]  # 40 copies of rows 0..7
    x = corpus[:8].astype(np.float32)
    ix = _check(corpus, x, 25)
    assert ix.last_stats()["uncertified"] >= 1


def test_hundreds_of_uncertified_queries_in_one_batch_also_asynchronously():
    """A corpus full of exact duplicates: more than 256 queries of ONE batch need the exact path
    (round 1 capped that at 256 and returned LXG_ETIES; with device outputs nobody even read the
    count).  Every query gets its own exact list now - host and asynchronous device results."""
    base, copies = 320, 80  # more copies than k' = 64: the cut falls inside the run of equal scores
    corpus = make_corpus(30000, 64)
    for j in range(base):
        corpus[base + j * copies : base + (j + 1) * copies] = corpus[j]  # 80 more copies of rows 0..319
    x = corpus[:base].astype(np.float32)
    ix = _check(corpus, x, 25)
    assert ix.last_stats()["uncertified"] > 256
    D, I = ix.search(x, 25, normalize=True)
    Dt, It = ix.search_torch(torch.from_numpy(x).cuda(), 25, normalize=True)
    torch.cuda.synchronize()
    assert np.array_equal(It.cpu().numpy(), I) and np.array_equal(Dt.cpu().numpy(), D)


def test_corpus_larger_than_l2_with_several_readers_per_slice():
    """230 MB corpus, 3 query blocks: the readers of a corpus slice are kept in step through the
    progress counters (ScanParams.progress, only enabled beyond the L2's size) - results unchanged."""
    corpus = make_corpus(300000, 384)
    ix = _check(corpus, make_queries(300, 384), 10)
    st = ix.last_stats()
    assert st["query_blocks"] == 3 and st["slices"] >= 30


def test_pinned_host_queries_are_read_in_place():
    """Page-locked queries (and the page-locked result arrays GpuIndexFlatIP.search allocates) are
    read / written by the kernels directly over PCIe - same bits as the staged path."""
    corpus = make_corpus(30000, 384)
    ix = _index(corpus)
    x = make_queries(300, 384)
    D, I = ix.search(x, 50, normalize=True)                       # pageable queries: staged copy
    xp = torch.from_numpy(x).pin_memory().numpy()
    Dp, Ip = ix.search(xp, 50, normalize=True)                    # pinned queries: zero copy
    assert np.array_equal(I, Ip) and np.array_equal(D, Dp)
    assert np.array_equal(xp, x)                                  # queries are not modified
    _check(corpus, xp, 50)


def test_more_queries_than_one_launch_holds():
    """nq > 148 * 128: lxg_search splits the batch over several launches."""
    corpus = make_corpus(3000, 64)
    x = make_queries(19500, 64)
    _check(corpus, x, 5)


def test_concurrent_host_threads_share_one_index():
    """SURVEY.md section 8(b) threading: calls on one handle from several host threads (the MCP
    event loop plus its executor) are serialised inside the library and all return exact results."""
    import threading

    corpus = make_corpus(20000, 128)
    ix = _index(corpus)
    xs = [make_queries(50 + 7 * i, 128, seed=40 + i) for i in range(6)]
    want = [ix.search(x, 10, normalize=True) for x in xs]
    got = [None] * len(xs)

    def work(i):
        for _ in range(5):
            got[i] = ix.search(xs[i], 10, normalize=True)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(xs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for (Dw, Iw), (Dg, Ig) in zip(want, got):
        assert np.array_equal(Iw, Ig) and np.array_equal(Dw, Dg)


# ------------------------------------------------------------------ committed fixtures -> CUDA path
from golden_cases import case_names, load_case  # noqa: E402


@pytest.mark.parametrize("name", case_names())
def test_committed_golden_vectors(name):
    """tests/golden/flat_ip_golden.npz (frozen exact-arithmetic rankings, inputs re-derived from seeds
    and checked by SHA-256) against the CUDA path through the C ABI: ids identical, scores to 1e-3."""
    c = load_case(name)
    ix = _index(c["corpus"])
    D, I = ix.search(c["x"], c["k"], normalize=c["normalize"])
    assert np.array_equal(I, c["I"])
    live = c["I"] >= 0
    mag = max(1.0, float(np.abs(c["D"][live]).max())) if live.any() else 1.0
    assert np.abs(D[live] - c["D"][live]).max() < 1e-3 * mag
    assert (D[~live] == ff.NEG_FLT_MAX).all()


@pytest.mark.parametrize("n,d,k", [(20000, 1024, 1000), (50000, 384, 50), (3000, 768, 10)])
def test_single_query_against_the_faiss_seq_path(n, d, k):
    """The reference's real request is nq = 1 (engine.py:237-250), which FAISS answers on its per-pair
    SIMD path (exhaustive_inner_product_seq: fp32 dot products in a fixed 8-lane order + CMin heap),
    restated deterministically in oracle/flat_ip.c lxo_knn_inner_product_seq.  The CUDA result must
    equal it wherever fp32 rounding cannot flip the order: ids may differ only at positions whose exact
    scores are closer than the fp32 accumulation noise, and then only as a permutation of the same ids."""
    from oracle import c_oracle

    corpus = make_corpus(n, d, dtype=np.float32)
    x = make_queries(1, d, seed=11)
    ix = _index(corpus)
    D, I = ix.search(x, k, normalize=True)
    xn = x.copy()
    c_oracle.renorm_l2(xn)  # FAISS' own fp32-order renorm
    Ds, Is = c_oracle.knn_inner_product_seq(xn, corpus, k)
    assert np.abs(D - Ds).max() < 1e-3
    D64, I64 = ff.flat_ip_search_f64(corpus, xn, k)
    differ = I != Is
    if differ.any():
        assert ff.ambiguous_positions(D64, tol=4e-6)[differ].all(), "differs from FAISS' seq path away from a near-tie"
        # same candidate set except for the boundary rank
        assert len(set(I[0]) ^ set(Is[0])) <= 2
    assert differ.mean() < 0.02


def test_async_searches_on_two_streams_share_one_index():
    """device-output searches return before their kernels finish; two of them on different
    streams used to race on the index workspaces.  The library now orders searches of one handle on
    the device, so interleaved async calls give the synchronous answers."""
    corpus = make_corpus(60000, 384)
    ix = _index(corpus)
    xs = [torch.from_numpy(make_queries(300, 384, seed=70 + i)).cuda() for i in range(6)]
    want = [ix.search(x.cpu().numpy(), 50, normalize=True) for x in xs]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    got = []
    for rep in range(3):
        got = []
        for i, x in enumerate(xs):
            with torch.cuda.stream(streams[i % 2]):
                got.append(ix.search_torch(x, 50, normalize=True))
        torch.cuda.synchronize()
        for (Dw, Iw), (Dg, Ig) in zip(want, got):
            assert np.array_equal(Iw, Ig.cpu().numpy()) and np.array_equal(Dw, Dg.cpu().numpy())
    assert ix.sync() == 0


def test_sync_reports_what_an_async_search_cannot():
    """lxg_index_sync: the number of queries the exact path re-did, for callers of the device-output form."""
    corpus = make_corpus(30000, 128)
    corpus[20000:20100] = corpus[17]
    x = torch.from_numpy(corpus[[17, 3]].astype(np.float32)).cuda()
    ix = _index(corpus)
    ix.search_torch(x, 20, normalize=True)
    assert ix.sync() >= 1


def test_index_on_second_device_from_a_worker_thread():
    """executor threads start on device 0; every entry point must switch to its handle's
    device.  Needs two GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import threading

    from lean_explore_b200 import GpuIndexFlatIP

    corpus = make_corpus(20000, 256)
    x = make_queries(40, 256)
    ix0 = _index(corpus)
    ix1 = GpuIndexFlatIP.from_tensor(torch.from_numpy(corpus).to("cuda:1"))
    want = ix0.search(x, 10, normalize=True)
    got = {}

    def work():
        got["r"] = ix1.search(x, 10, normalize=True)  # this thread's current device is 0

    t = threading.Thread(target=work)
    t.start()
    t.join()
    assert np.array_equal(want[1], got["r"][1]) and np.array_equal(want[0], got["r"][0])
    assert np.array_equal(ix0.search(x, 10, normalize=True)[1], want[1])
