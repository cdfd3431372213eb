"""CPU suite: host-side logic around the kernels - corpus codecs and loaders, the semantic-
retrieval mirror (with an injected index, the way the reference's own tests inject mocks,
tests/conftest.py:138-181 there), sharding arithmetic."""

import asyncio
import json
import sqlite3

from pathlib import Path

import numpy as np
import pytest

from conftest import make_corpus
from lean_explore_b200 import corpus as lc
from lean_explore_b200.semantic import SemanticRetriever
from lean_explore_b200.sharded import merge_topk_host, shard_rows
from oracle import faiss_flat as ff


# ------------------------------------------------------------------ BinaryEmbedding codec
def test_binary_embedding_round_trip():
    """reference tests/models/search_db_test.py:16-75: pack/unpack to 1e-6, None passthrough."""
    v = [0.1, -0.2, 3.5, 1e-7, 12345.678]
    blob = lc.pack_embedding(v)
    assert isinstance(blob, bytes) and len(blob) == 4 * len(v)
    back = lc.unpack_embedding(blob)
    assert all(abs(a - b) <= 1e-6 * max(1, abs(a)) for a, b in zip(v, back))
    assert lc.pack_embedding(None) is None and lc.unpack_embedding(None) is None
    assert lc.unpack_embedding(b"") == []
    big = list(np.random.default_rng(0).standard_normal(768).astype(np.float32))
    assert lc.unpack_embedding(lc.pack_embedding(big)) == [float(x) for x in big]


def _make_db(path, ids, mat, null_ids=()):
    con = sqlite3.connect(path)
    con.execute("CREATE TABLE declarations (id INTEGER PRIMARY KEY, name TEXT, informalization_embedding BLOB)")
    for i, row in zip(ids, mat):
        con.execute("INSERT INTO declarations VALUES (?, ?, ?)", (i, f"decl{i}", lc.pack_embedding(list(row))))
    for i in null_ids:
        con.execute("INSERT INTO declarations VALUES (?, ?, NULL)", (i, f"decl{i}"))
    con.commit()
    con.close()


def test_load_embeddings_from_database(tmp_path):
    """reference extract/index.py:45-78 + tests/extract/index_test.py:358-409: matrix rows in
    DB order, id list equals the DB ids, NULL embeddings skipped."""
    mat = make_corpus(300, 48, dtype=np.float32)
    ids = list(range(1, 301))
    db = tmp_path / "lean_explore.db"
    _make_db(db, ids, mat, null_ids=[1000, 1001])
    got_ids, got = lc.load_embeddings_from_database(db)
    assert got_ids == ids
    assert got.dtype == np.float32 and np.array_equal(got, mat)
    got_ids2, _ = lc.load_embeddings_from_database(f"sqlite+aiosqlite:///{db}")
    assert got_ids2 == ids


def test_load_embeddings_from_empty_database(tmp_path):
    db = tmp_path / "e.db"
    _make_db(db, [], np.zeros((0, 4), np.float32), null_ids=[1])
    ids, arr = lc.load_embeddings_from_database(db)
    assert ids == [] and arr.size == 0


# ------------------------------------------------------------------ FAISS index files
def test_flat_index_file_round_trip(tmp_path):
    mat = make_corpus(257, 40, dtype=np.float32)
    p = tmp_path / "flat.index"
    lc.write_flat_index(p, mat)
    got, info = lc.read_index_matrix(p)
    assert info["kind"] == "IxFI" and info["d"] == 40 and info["ntotal"] == 257 and info["metric_type"] == 0
    assert np.array_equal(got, mat)


@pytest.mark.parametrize("nlist,sparse", [(16, False), (256, None), (256, False)])
def test_ivfflat_index_file_reconstructs_label_order(tmp_path, nlist, sparse):
    """The shipped artefact is an IndexIVFFlat (extract/index.py:103-104): vectors are stored
    per inverted list with their labels; the reader must put row i back at label i."""
    rng = np.random.default_rng(5)
    mat = make_corpus(300, 24, dtype=np.float32)
    cent = make_corpus(nlist, 24, seed=9, dtype=np.float32)
    assign = rng.integers(0, min(nlist, 40), size=300)  # most lists empty when nlist = 256 ("sprs")
    p = tmp_path / "informalization_faiss.index"
    lc.write_ivfflat_index(p, mat, assign, cent, nprobe=1, sparse_sizes=sparse)
    got, info = lc.read_index_matrix(p)
    assert info["kind"] == "IwFl" and info["nlist"] == nlist and info["ntotal"] == 300
    assert np.array_equal(got, mat)


@pytest.mark.parametrize("name,rows,kind,nprobe", [("faiss_ivfflat_23x8.index", 23, "IwFl", 64),
                                                   ("faiss_ivfflat_sparse_7x8.index", 7, "IwFl", 1),
                                                   ("faiss_flat_23x8.index", 23, "IxFI", None)])
def test_committed_faiss_file_bytes_written_independently_of_this_package(name, rows, kind, nprobe):
    """tests/golden/*.index were assembled field by field from FAISS' published serialisation order
    (faiss/impl/index_write.cpp) by tests/golden/make_faiss_file_golden.py, which does not import this
    package: the reader is no longer checked only against its own writer.  They exercise an Array
    direct map with entries, "full" and "sprs" list sizes, empty lists and unsorted labels in a list."""
    golden = Path(__file__).resolve().parent / "golden"
    want = np.load(golden / "faiss_file_vectors_23x8.npy")[:rows]
    got, info = lc.read_index_matrix(golden / name)
    assert info["kind"] == kind and info["ntotal"] == rows and info["d"] == 8 and info["metric_type"] == 0
    assert info.get("nprobe") == nprobe
    assert got.dtype == np.float32 and np.array_equal(got, want)


def test_index_file_errors(tmp_path):
    p = tmp_path / "x.index"
    p.write_bytes(b"")
    with pytest.raises(ValueError):
        lc.read_index_matrix(p)
    p.write_bytes(b"IHNf" + b"\0" * 64)
    with pytest.raises(ValueError, match="unsupported"):
        lc.read_index_matrix(p)
    mat = make_corpus(10, 8, dtype=np.float32)
    lc.write_flat_index(p, mat)
    p.write_bytes(p.read_bytes()[:-5])
    with pytest.raises(ValueError, match="truncated"):
        lc.read_index_matrix(p)


# ------------------------------------------------------------------ semantic retrieval mirror
class _OracleIndex:
    """Stands in for the GPU index on the CPU suite (same duck type; test-only)."""

    def __init__(self, mat):
        self.mat = mat
        self.d = mat.shape[1]
        self.ntotal = mat.shape[0]
        self.nprobe = 1

    def search(self, x, k, normalize=False):
        x = np.array(x, dtype=np.float32)
        if normalize:
            ff.normalize_L2(x)
        return ff.flat_ip_search(self.mat, x, k)


class _FakeEmbeddingClient:
    model_name = "fake"

    def __init__(self, table):
        self.table = table
        self.calls = []

    async def embed(self, texts, is_query=False):
        from lean_explore_b200.embedding_client import EmbeddingResponse

        self.calls.append((list(texts), is_query))
        return EmbeddingResponse(texts=list(texts), embeddings=[self.table[t] for t in texts], model="fake")


def _retriever(tmp_path, mat, id_map, client):
    ip = tmp_path / "informalization_faiss.index"
    mp = tmp_path / "informalization_faiss_ids_map.json"
    lc.write_flat_index(ip, mat)
    mp.write_text(json.dumps(id_map))
    r = SemanticRetriever(embedding_client=client, base_path=tmp_path)
    r._faiss_informal_index = _OracleIndex(mat)       # as reference tests do: engine._faiss_informal_index = mock
    r._faiss_informal_id_map = lc.load_ids_map(mp)
    return r


def test_missing_artefacts_raise_file_not_found(tmp_path):
    """reference tests/search/engine_test.py:190-201."""
    with pytest.raises(FileNotFoundError, match="Required file not found at .*lean-explore data fetch"):
        SemanticRetriever(embedding_client=object(), base_path=tmp_path)


def test_retrieve_semantic_candidates_single_and_batch(tmp_path):
    mat = make_corpus(400, 32, dtype=np.float32)
    id_map = [1000 + (i // 2) for i in range(400)]  # two rows per declaration -> max-dedupe matters
    table = {"q0": [float(v) for v in mat[10] * 3], "q1": [float(v) for v in mat[77] * 0.5]}
    client = _FakeEmbeddingClient(table)
    r = _retriever(tmp_path, mat, id_map, client)
    got = asyncio.run(r._retrieve_semantic_candidates("q0", 20))
    assert client.calls[-1] == (["q0"], True)  # is_query=True, engine.py:237
    assert r.faiss_informal_index.nprobe == 64  # engine.py:247-248
    want = ff.retrieve_semantic_candidates(ff.IndexFlatIP(32), id_map, table["q0"], 20)  # empty index -> {}
    assert want == {}
    ix = ff.IndexFlatIP(32)
    ix.add(mat)
    want = ff.retrieve_semantic_candidates(ix, id_map, table["q0"], 20)
    assert got.keys() == want.keys() and max(got, key=got.get) == 1005
    assert all(abs(got[k] - want[k]) < 1e-6 for k in got)
    batch = asyncio.run(r.retrieve_semantic_candidates_batch(["q0", "q1"], 20))
    assert batch[0].keys() == got.keys() and all(abs(batch[0][k] - got[k]) < 1e-6 for k in got)
    assert max(batch[1], key=batch[1].get) == 1000 + 77 // 2
    assert asyncio.run(r.retrieve_semantic_candidates_batch([], 5)) == []

    class _ArrayClient(_FakeEmbeddingClient):  # GpuEmbeddingClient's batched twin: numpy, no per-float lists
        def embed_array(self, texts, is_query=False):
            self.calls.append((list(texts), is_query, "array"))
            return np.array([self.table[t] for t in texts], dtype=np.float32)

    fast = _ArrayClient(table)
    r2 = _retriever(tmp_path, mat, id_map, fast)
    batch2 = asyncio.run(r2.retrieve_semantic_candidates_batch(["q0", "q1"], 20))
    assert fast.calls == [(["q0", "q1"], True, "array")]
    assert [sorted(b.items()) for b in batch2] == [sorted(b.items()) for b in batch]


# ------------------------------------------------------------------ sharding arithmetic
def test_shard_rows_partition():
    for n, w in [(10, 3), (16_000_000, 8), (5, 8), (0, 2)]:
        spans = [shard_rows(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert all(hi >= lo for lo, hi in spans)


def test_merge_topk_host_equals_single_search():
    mat = make_corpus(1000, 16, dtype=np.float32)
    x = make_corpus(6, 16, seed=3, dtype=np.float32)
    k = 9
    Df, If = ff.flat_ip_search_f64(mat, x, k)
    dg, ig = [], []
    for r in range(3):
        lo, hi = shard_rows(1000, 3, r)
        d, i = ff.flat_ip_search_f64(mat[lo:hi], x, k)
        dg.append(d)
        ig.append(np.where(i >= 0, i + lo, -1))
    D, I = merge_topk_host(np.stack(dg), np.stack(ig), k)
    assert np.array_equal(I, If) and np.allclose(D, Df.astype(np.float32))
