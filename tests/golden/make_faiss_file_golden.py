"""Hand-assembles the bytes `faiss.write_index` emits for the index the reference builds
(`extract/index.py:103-116`: IndexIVFFlat over an IndexFlatIP quantizer, METRIC_INNER_PRODUCT) and for a
plain IndexFlatIP, field by field from FAISS' serialisation routine (faiss/impl/index_write.cpp,
FAISS 1.7/1.8 - the range `faiss-cpu>=1.7` of the reference's pyproject.toml:36 resolves to):

  write_index_header : int d | idx_t ntotal | idx_t dummy (1 << 20) x 2 | bool is_trained | int metric_type
                       (| float metric_arg when metric_type > 1)
  IndexFlat          : fourcc "IxFI" (inner product) / "IxF2" (L2) | header | WRITEXBVECTOR(codes):
                       size_t count-of-4-byte-words | raw fp32
  write_ivf_header   : header | size_t nlist | size_t nprobe | write_index(quantizer) | write_direct_map
  write_direct_map   : char type (0 NoMap, 1 Array, 2 Hashtable) | WRITEVECTOR(array: idx_t)
                       (| WRITEVECTOR(pairs) for Hashtable)
  IndexIVFFlat       : fourcc "IwFl" | ivf header | write_InvertedLists
  ArrayInvertedLists : fourcc "ilar" | size_t nlist | size_t code_size | fourcc "full" + WRITEVECTOR(sizes)
                       when more than nlist / 2 lists are non-empty, else fourcc "sprs" +
                       WRITEVECTOR([list no, size, list no, size, ...]) | for every non-empty list:
                       codes (size * code_size bytes) then ids (size * idx_t)
  WRITEVECTOR        : size_t count | raw elements.   All little endian, size_t / idx_t are 8 bytes.

faiss itself is not installable here (no network), so this is NOT output of faiss - "format pinned to the
published serialisation order, not to a FAISS-written file" (DESIGN.md).  It deliberately does not import
lean_explore_b200: the reader (corpus.read_index_matrix) must not be tested only against its own writer.
Differences from the repo's writer that this fixture exercises: an Array direct map with entries
(type 1), is_trained / dummy fields written literally, the quantizer's own header, a metric_arg-free
inner-product metric, lists visited in list order with ids NOT sorted inside a list.

Run: python tests/golden/make_faiss_file_golden.py   (rewrites the two .index files and the .npy)
"""
from __future__ import annotations

import io
import struct
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
METRIC_INNER_PRODUCT = 0


def u32(tag: str) -> bytes:
    assert len(tag) == 4
    return struct.pack("<I", sum(ord(c) << (8 * i) for i, c in enumerate(tag)))


def header(d: int, ntotal: int, metric: int) -> bytes:
    out = struct.pack("<i", d) + struct.pack("<q", ntotal)
    out += struct.pack("<q", 1 << 20) + struct.pack("<q", 1 << 20)
    out += struct.pack("<?", True) + struct.pack("<i", metric)
    return out


def flat(vectors: np.ndarray, metric: int) -> bytes:
    raw = vectors.astype("<f4").tobytes()
    return u32("IxFI" if metric == METRIC_INNER_PRODUCT else "IxF2") + header(vectors.shape[1], vectors.shape[0], metric) + \
        struct.pack("<Q", len(raw) // 4) + raw


def main() -> None:
    rng = np.random.default_rng(20260101)
    n, d, nlist, nprobe = 23, 8, 6, 64
    x = rng.standard_normal((n, d)).astype(np.float32)
    centroids = rng.standard_normal((nlist, d)).astype(np.float32)
    # list membership in insertion order per list, as IndexIVF::add appends (list 1 and 4 stay empty)
    lists = {0: [3, 0, 17, 9], 2: [1, 22, 2, 14, 5, 8], 3: [4, 21, 6, 7, 20, 10], 5: [19, 11, 12, 13, 15, 16, 18]}
    assert sorted(i for l in lists.values() for i in l) == list(range(n))

    buf = io.BytesIO()
    buf.write(u32("IwFl"))
    buf.write(header(d, n, METRIC_INNER_PRODUCT))
    buf.write(struct.pack("<Q", nlist))
    buf.write(struct.pack("<Q", nprobe))
    buf.write(flat(centroids, METRIC_INNER_PRODUCT))            # the coarse quantizer, a full IndexFlatIP
    # direct map of type Array: array[label] = (list_no << 32) | offset
    lo = {}
    for ln, members in lists.items():
        for off, lab in enumerate(members):
            lo[lab] = (ln << 32) | off
    buf.write(struct.pack("<b", 1))
    buf.write(struct.pack("<Q", n))
    buf.write(b"".join(struct.pack("<q", lo[i]) for i in range(n)))
    buf.write(u32("ilar"))
    buf.write(struct.pack("<Q", nlist))
    buf.write(struct.pack("<Q", 4 * d))
    buf.write(u32("full"))                                      # 4 of 6 lists are non-empty (> nlist / 2)
    buf.write(struct.pack("<Q", nlist))
    for ln in range(nlist):
        buf.write(struct.pack("<Q", len(lists.get(ln, []))))
    for ln in range(nlist):
        members = lists.get(ln, [])
        if members:
            buf.write(b"".join(x[i].astype("<f4").tobytes() for i in members))
            buf.write(b"".join(struct.pack("<q", i) for i in members))
    (HERE / "faiss_ivfflat_23x8.index").write_bytes(buf.getvalue())

    # sparse list sizes: 2 of 6 lists non-empty (<= nlist / 2), NoMap direct map
    lists2 = {1: [5, 1, 0, 6], 4: [2, 4, 3]}
    buf = io.BytesIO()
    buf.write(u32("IwFl") + header(d, 7, METRIC_INNER_PRODUCT) + struct.pack("<QQ", nlist, 1))
    buf.write(flat(centroids, METRIC_INNER_PRODUCT))
    buf.write(struct.pack("<b", 0) + struct.pack("<Q", 0))
    buf.write(u32("ilar") + struct.pack("<QQ", nlist, 4 * d) + u32("sprs"))
    buf.write(struct.pack("<Q", 4) + struct.pack("<QQQQ", 1, 4, 4, 3))
    for ln in (1, 4):
        buf.write(b"".join(x[i].astype("<f4").tobytes() for i in lists2[ln]))
        buf.write(b"".join(struct.pack("<q", i) for i in lists2[ln]))
    (HERE / "faiss_ivfflat_sparse_7x8.index").write_bytes(buf.getvalue())

    (HERE / "faiss_flat_23x8.index").write_bytes(flat(x, METRIC_INNER_PRODUCT))
    np.save(HERE / "faiss_file_vectors_23x8.npy", x)


if __name__ == "__main__":
    main()
