"""Corpus ingestion: how the declaration-embedding matrix reaches HBM.

Host-side byte parsing only (no arithmetic).  Mirrors the reference's definitions of the
corpus matrix and its artefacts:

* ``BinaryEmbedding`` row codec - ``src/lean_explore/models/search_db.py:14-35`` and
  ``src/lean_explore/extract/embeddings.py:94-104``: native-endian fp32, 4*d bytes per row.
* ``_load_embeddings_from_database`` - ``src/lean_explore/extract/index.py:45-78``:
  ``SELECT id, informalization_embedding FROM declarations WHERE ... IS NOT NULL`` ->
  ``(declaration_ids, float32 [N, d])``; row i of the matrix <-> declaration_ids[i].
* ``faiss.write_index`` / ``faiss.read_index`` artefacts - ``extract/index.py:171-182`` and
  ``search/engine.py:151-161``: ``informalization_faiss.index`` (an ``IndexIVFFlat``, fourcc
  "IwFl", or a flat index "IxFI"/"IxF2"/"IxFl") plus ``informalization_faiss_ids_map.json``.
  The reader reconstructs the [N, d] fp32 matrix in LABEL order (FAISS labels are the
  sequential row numbers, there is no IndexIDMap), which is all a brute-force index needs.

The on-disk layout of FAISS indexes is restated from faiss/impl/index_write.cpp and
index_read.cpp (faiss 1.7 - 1.12; the wheel is not installable here, so the reader is tested
by round-tripping files produced by the writer below, which follows the same source).
"""

from __future__ import annotations

import json
import sqlite3
import struct
from pathlib import Path

import numpy as np

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1


# ----------------------------------------------------------------------------- row codec
def pack_embedding(value) -> bytes | None:
    """``BinaryEmbedding.process_bind_param`` (search_db.py:24-28)."""
    if value is None:
        return None
    return struct.pack(f"{len(value)}f", *value)


def unpack_embedding(value: bytes | None) -> list[float] | None:
    """``BinaryEmbedding.process_result_value`` (search_db.py:30-35)."""
    if value is None:
        return None
    num_floats = len(value) // 4
    return list(struct.unpack(f"{num_floats}f", value))


def load_embeddings_from_database(db_path, embedding_field: str = "informalization_embedding"):
    """``_load_embeddings_from_database`` (extract/index.py:45-78) over the stdlib sqlite3
    driver: returns ``(declaration_ids: list[int], embeddings: float32 [N, d])`` in the order
    SQLite yields the rows, or ``([], np.array([]))`` when no row has an embedding."""
    if not embedding_field.isidentifier():
        raise ValueError(f"bad column name {embedding_field!r}")
    path = str(db_path)
    for prefix in ("sqlite+aiosqlite:///", "sqlite:///"):
        if path.startswith(prefix):
            path = path[len(prefix):]
    con = sqlite3.connect(f"file:{path}?mode=ro", uri=True)
    try:
        rows = con.execute(
            f"SELECT id, {embedding_field} FROM declarations WHERE {embedding_field} IS NOT NULL"
        ).fetchall()
    finally:
        con.close()
    if not rows:
        return [], np.array([])
    ids = [int(r[0]) for r in rows]
    d = len(rows[0][1]) // 4
    out = np.empty((len(rows), d), dtype=np.float32)
    for i, (_, blob) in enumerate(rows):
        if len(blob) != 4 * d:
            raise ValueError(f"row {ids[i]}: embedding has {len(blob) // 4} dims, expected {d}")
        out[i] = np.frombuffer(blob, dtype=np.float32)
    return ids, out


def load_ids_map(path) -> list[int]:
    """``informalization_faiss_ids_map.json`` (engine.py:160-161): a JSON list, row i -> id."""
    with open(path) as f:
        ids = json.load(f)
    if not isinstance(ids, list):
        raise ValueError(f"{path}: id map must be a JSON list")
    return ids


# ----------------------------------------------------------------------------- FAISS files
class _Reader:
    def __init__(self, data: bytes):
        self.b = memoryview(data)
        self.o = 0

    def take(self, n: int) -> memoryview:
        if self.o + n > len(self.b):
            raise ValueError("truncated FAISS index file")
        v = self.b[self.o : self.o + n]
        self.o += n
        return v

    def fourcc(self) -> str:
        return bytes(self.take(4)).decode("latin1")

    def scalar(self, fmt: str):
        return struct.unpack("<" + fmt, self.take(struct.calcsize("<" + fmt)))[0]

    def vector(self, dtype) -> np.ndarray:
        n = self.scalar("Q")
        dt = np.dtype(dtype)
        return np.frombuffer(self.take(n * dt.itemsize), dtype=dt)


def _read_header(r: _Reader) -> dict:
    # write_index_header: int d; idx_t ntotal; idx_t dummy x2; bool is_trained; int metric_type
    h = {"d": r.scalar("i"), "ntotal": r.scalar("q")}
    r.scalar("q")
    r.scalar("q")
    h["is_trained"] = bool(r.scalar("B"))
    h["metric_type"] = r.scalar("i")
    if h["metric_type"] > 1:
        h["metric_arg"] = r.scalar("f")
    return h


def _read_flat(r: _Reader, fourcc: str) -> tuple[dict, np.ndarray]:
    h = _read_header(r)
    xb = r.vector(np.float32)  # WRITEXBVECTOR: count of floats, then the floats
    if xb.size != h["ntotal"] * h["d"]:
        raise ValueError("flat index: payload size does not match ntotal * d")
    return h, xb.reshape(h["ntotal"], h["d"])


def read_index_matrix(path) -> tuple[np.ndarray, dict]:
    """Parse a FAISS index file and return ``(float32 [ntotal, d] in label order, info)``.

    Supports the two kinds the reference can produce or a user can substitute: flat
    ("IxFI", "IxF2", "IxFl") and ``IndexIVFFlat`` ("IwFl") with array inverted lists.
    """
    data = Path(path).read_bytes()
    if not data:
        raise ValueError(f"{path}: empty FAISS index file")
    r = _Reader(data)
    fourcc = r.fourcc()
    if fourcc in ("IxFI", "IxF2", "IxFl"):
        h, m = _read_flat(r, fourcc)
        return np.array(m, dtype=np.float32), {"kind": fourcc, **h}
    if fourcc != "IwFl":
        raise ValueError(f"{path}: unsupported FAISS index type {fourcc!r}")
    h = _read_header(r)
    nlist = r.scalar("Q")
    nprobe = r.scalar("Q")
    qcc = r.fourcc()
    if qcc not in ("IxFI", "IxF2", "IxFl"):
        raise ValueError(f"coarse quantizer {qcc!r} not supported")
    _read_flat(r, qcc)
    # direct map: char type; vector<idx_t> array; (+ hashtable pairs when type == 2)
    dm_type = r.scalar("b")
    r.vector(np.int64)
    if dm_type == 2:
        r.vector(np.dtype([("a", "<i8"), ("b", "<i8")]))
    il = r.fourcc()
    if il == "il00":
        raise ValueError("index has no inverted lists")
    if il != "ilar":
        raise ValueError(f"inverted lists {il!r} not supported")
    il_nlist = r.scalar("Q")
    code_size = r.scalar("Q")
    d, n = h["d"], h["ntotal"]
    if code_size != 4 * d or il_nlist != nlist:
        raise ValueError("IVFFlat: code_size / nlist mismatch")
    fmt = r.fourcc()
    sizes = np.zeros(nlist, dtype=np.int64)
    if fmt == "full":
        sizes[:] = r.vector(np.uint64).astype(np.int64)
    elif fmt == "sprs":
        pairs = r.vector(np.uint64).astype(np.int64).reshape(-1, 2)
        sizes[pairs[:, 0]] = pairs[:, 1]
    else:
        raise ValueError(f"list-size format {fmt!r} not supported")
    out = np.zeros((n, d), dtype=np.float32)
    seen = np.zeros(n, dtype=bool)
    for ln in sizes:
        ln = int(ln)
        if ln == 0:
            continue
        codes = np.frombuffer(r.take(ln * code_size), dtype=np.float32).reshape(ln, d)
        ids = np.frombuffer(r.take(ln * 8), dtype=np.int64)
        if ids.min() < 0 or ids.max() >= n:
            raise ValueError("IVFFlat: label out of range (IndexIDMap-style labels are not supported)")
        out[ids] = codes
        seen[ids] = True
    if not seen.all():
        raise ValueError("IVFFlat: some labels have no stored vector")
    return out, {"kind": fourcc, "nlist": int(nlist), "nprobe": int(nprobe), **h}


def _write_header(f, d: int, ntotal: int, metric: int) -> None:
    f.write(struct.pack("<iqqqBi", d, ntotal, 1 << 20, 1 << 20, 1, metric))


def write_flat_index(path, matrix: np.ndarray, metric: int = METRIC_INNER_PRODUCT) -> None:
    """``faiss.write_index(IndexFlatIP)`` layout (test fixture / export helper)."""
    m = np.ascontiguousarray(matrix, dtype=np.float32)
    with open(path, "wb") as f:
        f.write(b"IxFI" if metric == METRIC_INNER_PRODUCT else b"IxF2")
        _write_header(f, m.shape[1], m.shape[0], metric)
        f.write(struct.pack("<Q", m.size))
        f.write(m.tobytes())


def write_ivfflat_index(path, matrix: np.ndarray, assign: np.ndarray, centroids: np.ndarray,
                        nprobe: int = 1, sparse_sizes: bool | None = None) -> None:
    """``faiss.write_index(IndexIVFFlat)`` layout for a given list assignment (test fixture:
    what ``extract/index.py:103-116`` would serialise after train + add)."""
    m = np.ascontiguousarray(matrix, dtype=np.float32)
    c = np.ascontiguousarray(centroids, dtype=np.float32)
    n, d = m.shape
    nlist = c.shape[0]
    lists = [np.flatnonzero(assign == i) for i in range(nlist)]
    nonzero = sum(1 for l in lists if len(l))
    if sparse_sizes is None:
        sparse_sizes = nonzero * 2 < nlist  # faiss picks "sprs" when most lists are empty
    with open(path, "wb") as f:
        f.write(b"IwFl")
        _write_header(f, d, n, METRIC_INNER_PRODUCT)
        f.write(struct.pack("<QQ", nlist, nprobe))
        f.write(b"IxFI")
        _write_header(f, d, nlist, METRIC_INNER_PRODUCT)
        f.write(struct.pack("<Q", c.size))
        f.write(c.tobytes())
        f.write(struct.pack("<bQ", 0, 0))  # direct map: NoMap, empty array
        f.write(b"ilar")
        f.write(struct.pack("<QQ", nlist, 4 * d))
        if sparse_sizes:
            f.write(b"sprs")
            pairs = [(i, len(l)) for i, l in enumerate(lists) if len(l)]
            f.write(struct.pack("<Q", 2 * len(pairs)))
            for i, ln in pairs:
                f.write(struct.pack("<QQ", i, ln))
        else:
            f.write(b"full")
            f.write(struct.pack("<Q", nlist))
            f.write(np.array([len(l) for l in lists], dtype=np.uint64).tobytes())
        for l in lists:
            if len(l):
                f.write(m[l].tobytes())
                f.write(l.astype(np.int64).tobytes())
