"""FAISS on-disk format of the IVFFlat index (recommendit_b200/faiss_io.py) — CPU tests.

faiss itself is absent here, so the byte layout is pinned against an INDEPENDENT restatement: the expected bytes of a tiny index are
assembled below field by field with struct.pack following faiss/impl/index_write.cpp (write_index_header, write_ivf_header,
write_direct_map, write_InvertedLists), not by calling the writer.  The cross-check against the real library lives in
tests/test_gpu_ivf.py::test_real_faiss_cross_check (skipped when `import faiss` fails)."""
import io
import struct

import numpy as np
import pytest

from recommendit_b200 import faiss_io as F


def tiny():
    d, nlist = 2, 4
    cen = np.array([[1, 0], [0, 1], [-1, 0], [0, -1]], dtype=np.float32)
    sizes = [2, 0, 3, 1]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    vecs = (np.arange(12, dtype=np.float32).reshape(6, 2) + 0.5)
    ids = np.array([0, 4, 1, 2, 5, 3], dtype=np.int64)
    return F.IVFFlatData(d=d, nlist=nlist, nprobe=3, metric_type=F.METRIC_INNER_PRODUCT, centroids=cen, offsets=off, list_ids=ids, list_vecs=vecs)


def expected_bytes(x, sparse):
    hdr = lambda d, n: struct.pack("<i", d) + struct.pack("<q", n) + struct.pack("<q", 1 << 20) * 2 + struct.pack("<?", True) + struct.pack("<i", 0)
    b = b"IwFl" + hdr(x.d, x.ntotal) + struct.pack("<Q", x.nlist) + struct.pack("<Q", x.nprobe)
    b += b"IxFI" + hdr(x.d, x.nlist) + struct.pack("<Q", x.nlist * x.d) + x.centroids.astype("<f4").tobytes()
    b += struct.pack("<b", 0) + struct.pack("<Q", 0)                         # direct map: NoMap + empty idx_t array
    b += b"ilar" + struct.pack("<Q", x.nlist) + struct.pack("<Q", 4 * x.d)
    sizes = np.diff(x.offsets)
    if sparse:
        nz = [(i, int(n)) for i, n in enumerate(sizes) if n]
        b += b"sprs" + struct.pack("<Q", 2 * len(nz)) + b"".join(struct.pack("<QQ", i, n) for i, n in nz)
    else:
        b += b"full" + struct.pack("<Q", x.nlist) + b"".join(struct.pack("<Q", int(n)) for n in sizes)
    for i in range(x.nlist):
        a, e = int(x.offsets[i]), int(x.offsets[i + 1])
        if e > a:
            b += x.list_vecs[a:e].astype("<f4").tobytes() + x.list_ids[a:e].astype("<i8").tobytes()
    return b


def test_writer_emits_the_faiss_layout_full_and_sparse():
    x = tiny()                                   # 3 of 4 lists non-empty (> nlist/2): "full"
    f = io.BytesIO(); F.write_ivfflat(f, x)
    assert f.getvalue() == expected_bytes(x, sparse=False)
    x2 = tiny()
    x2.offsets = np.array([0, 6, 6, 6, 6], dtype=np.int64)                   # 1 of 4 lists non-empty: "sprs"
    f = io.BytesIO(); F.write_ivfflat(f, x2)
    assert f.getvalue() == expected_bytes(x2, sparse=True)


@pytest.mark.parametrize("sparse", [False, True])
def test_reader_parses_the_independent_restatement(sparse):
    x = tiny()
    if sparse:
        x.offsets = np.array([0, 0, 6, 6, 6], dtype=np.int64)
    y = F.read_ivfflat(io.BytesIO(expected_bytes(x, sparse)))
    assert (y.d, y.nlist, y.nprobe, y.metric_type, y.ntotal, y.is_trained) == (2, 4, 3, 0, 6, True)
    for a, b in ((y.centroids, x.centroids), (y.offsets, x.offsets), (y.list_ids, x.list_ids), (y.list_vecs, x.list_vecs)):
        assert np.array_equal(a, b)


def test_reader_skips_direct_maps_and_rejects_other_indexes():
    x = tiny()
    raw = expected_bytes(x, False)
    dm = raw.index(b"ilar") - 9
    # array direct map (type 1) with ntotal entries
    with_map = raw[:dm] + struct.pack("<b", 1) + struct.pack("<Q", 6) + np.arange(6, dtype="<i8").tobytes() + raw[dm + 9:]
    assert np.array_equal(F.read_ivfflat(io.BytesIO(with_map)).list_ids, x.list_ids)
    # hash-table direct map (type 2): empty array + pairs
    with_hash = raw[:dm] + struct.pack("<b", 2) + struct.pack("<Q", 0) + struct.pack("<Q", 2) + np.arange(4, dtype="<i8").tobytes() + raw[dm + 9:]
    assert np.array_equal(F.read_ivfflat(io.BytesIO(with_hash)).list_vecs, x.list_vecs)
    with pytest.raises(F.FaissFormatError, match="IndexIVFFlat"):
        F.read_ivfflat(io.BytesIO(b"IxFI" + raw[4:]))
    with pytest.raises(F.FaissFormatError, match="end of file"):
        F.read_ivfflat(io.BytesIO(raw[:-3]))


def test_round_trip_random_index():
    rng = np.random.default_rng(0)
    d, nlist, n = 32, 10, 500
    sizes = rng.multinomial(n, np.ones(nlist) / nlist)
    x = F.IVFFlatData(d=d, nlist=nlist, nprobe=5, metric_type=0, centroids=rng.standard_normal((nlist, d)).astype(np.float32),
                      offsets=np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64), list_ids=rng.permutation(n).astype(np.int64),
                      list_vecs=rng.standard_normal((n, d)).astype(np.float32))
    f = io.BytesIO(); F.write_ivfflat(f, x); f.seek(0)
    y = F.read_ivfflat(f)
    assert np.array_equal(y.list_vecs, x.list_vecs) and np.array_equal(y.list_ids, x.list_ids) and np.array_equal(y.offsets, x.offsets)
    assert f.read() == b""                       # nothing left unread
