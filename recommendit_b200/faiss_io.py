"""FAISS on-disk interop for the IVFFlat index (second half): read and write the file
``faiss.write_index`` / ``faiss.read_index`` exchange for ``IndexIVFFlat(IndexFlatIP(d), d, nlist, METRIC_INNER_PRODUCT)`` —
the only index type the reference builds (``src/models/faiss_index.py:68-69``, written at ``:164``, read back at ``:196``).

Pure NumPy / ``struct``: no FAISS, no CUDA — the layout below restates ``faiss/impl/index_write.cpp`` (faiss 1.7.x / 1.8.x;
``faiss-cpu>=1.7.4`` in the reference's requirements.txt:2).  All integers little-endian::

    "IwFl"                                              fourcc of IndexIVFFlat
    index header      d:i32  ntotal:i64  dummy:i64 (1<<20)  dummy:i64 (1<<20)  is_trained:u8  metric_type:i32 (0 = inner product, 1 = L2)
                      [metric_arg:f32 when metric_type > 1]
    nlist:u64  nprobe:u64
    quantizer         "IxFI" (IndexFlatIP; "IxF2" L2, "IxFl" other)  + index header (d, ntotal = nlist, …)  + n_floats:u64 + f32[nlist·d]
    direct map        type:u8 (0 = none, 1 = array, 2 = hash table)  + n:u64 + i64[n]   [+ n:u64 + (i64, i64)[n] for type 2]
    inverted lists    "ilar"  nlist:u64  code_size:u64 (= 4·d)
                      "full" + n:u64 + u64[nlist] list sizes      — or —   "sprs" + n:u64 + u64[n] = (list, size) pairs of non-empty lists
                      per non-empty list, in list order:  codes u8[size·code_size] (the raw f32 vectors)  then  ids i64[size]

FAISS itself is absent from this image, so the writer cannot be checked against ``faiss.read_index`` here;
``tests/test_faiss_io.py`` pins the byte layout against an independent ``struct.pack`` restatement, and
``tests/test_gpu_ivf.py::test_real_faiss_cross_check`` runs both directions against the real library whenever
``import faiss`` succeeds (skipped otherwise).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import BinaryIO, List, Optional

import numpy as np

METRIC_INNER_PRODUCT, METRIC_L2 = 0, 1
_DUMMY = 1 << 20


class FaissFormatError(ValueError):
    pass


@dataclass
class IVFFlatData:
    """Host image of an IndexIVFFlat: what both libraries need to search."""
    d: int
    nlist: int
    nprobe: int
    metric_type: int
    centroids: np.ndarray          # f32 [nlist, d]   (the IndexFlat quantizer's vectors)
    offsets: np.ndarray            # i64 [nlist + 1]  CSR over the lists
    list_ids: np.ndarray           # i64 [ntotal]     FAISS internal ids (sequential add order), list-contiguous
    list_vecs: np.ndarray          # f32 [ntotal, d]  list-contiguous, insertion order inside a list
    is_trained: bool = True

    @property
    def ntotal(self) -> int:
        return int(self.offsets[-1])


def _rd(f: BinaryIO, fmt: str):
    size = struct.calcsize(fmt)
    b = f.read(size)
    if len(b) != size:
        raise FaissFormatError("unexpected end of file")
    v = struct.unpack(fmt, b)
    return v[0] if len(v) == 1 else v


def _rd_array(f: BinaryIO, dtype, count: int) -> np.ndarray:
    nbytes = int(count) * np.dtype(dtype).itemsize
    b = f.read(nbytes)
    if len(b) != nbytes:
        raise FaissFormatError("unexpected end of file")
    return np.frombuffer(b, dtype=dtype, count=int(count)).copy()


def _read_header(f: BinaryIO):
    d = _rd(f, "<i")
    ntotal = _rd(f, "<q")
    _rd(f, "<q"); _rd(f, "<q")
    is_trained = bool(_rd(f, "<B"))
    metric = _rd(f, "<i")
    if metric > 1:
        _rd(f, "<f")
    return d, ntotal, is_trained, metric


def _write_header(f: BinaryIO, d: int, ntotal: int, is_trained: bool, metric: int) -> None:
    f.write(struct.pack("<iqqqBi", d, ntotal, _DUMMY, _DUMMY, 1 if is_trained else 0, metric))


def read_ivfflat(f: BinaryIO) -> IVFFlatData:
    """Parse a ``faiss.write_index`` file holding an IndexIVFFlat over an IndexFlat quantizer."""
    four = f.read(4)
    if four != b"IwFl":
        raise FaissFormatError(f"not an IndexIVFFlat file (fourcc {four!r}; this reader handles 'IwFl', the index the reference builds)")
    d, ntotal, is_trained, metric = _read_header(f)
    nlist, nprobe = _rd(f, "<Q"), _rd(f, "<Q")
    qfour = f.read(4)
    if qfour not in (b"IxFI", b"IxF2", b"IxFl"):
        raise FaissFormatError(f"unsupported coarse quantizer {qfour!r} (expected a flat index)")
    qd, qn, _, _ = _read_header(f)
    n_floats = _rd(f, "<Q")
    if qd != d or qn != nlist or n_floats != nlist * d:
        raise FaissFormatError(f"quantizer shape mismatch: d={qd}, ntotal={qn}, floats={n_floats} for nlist={nlist}, d={d}")
    centroids = _rd_array(f, "<f4", n_floats).reshape(nlist, d)
    dm_type = _rd(f, "<B")
    _rd_array(f, "<i8", _rd(f, "<Q"))
    if dm_type == 2:
        _rd_array(f, "<i8", 2 * _rd(f, "<Q"))
    il = f.read(4)
    if il != b"ilar":
        raise FaissFormatError(f"unsupported inverted-list container {il!r} (expected ArrayInvertedLists 'ilar')")
    il_nlist, code_size = _rd(f, "<Q"), _rd(f, "<Q")
    if il_nlist != nlist or code_size != 4 * d:
        raise FaissFormatError(f"inverted lists: nlist={il_nlist}, code_size={code_size} for nlist={nlist}, d={d}")
    kind = f.read(4)
    sizes = np.zeros(nlist, dtype=np.int64)
    if kind == b"full":
        v = _rd_array(f, "<u8", _rd(f, "<Q"))
        if len(v) != nlist:
            raise FaissFormatError("'full' size table does not have nlist entries")
        sizes[:] = v
    elif kind == b"sprs":
        v = _rd_array(f, "<u8", _rd(f, "<Q")).reshape(-1, 2)
        sizes[v[:, 0].astype(np.int64)] = v[:, 1]
    else:
        raise FaissFormatError(f"unknown list-size encoding {kind!r}")
    offsets = np.zeros(nlist + 1, dtype=np.int64)
    np.cumsum(sizes, out=offsets[1:])
    if offsets[-1] != ntotal:
        raise FaissFormatError(f"list sizes add up to {offsets[-1]}, header says ntotal={ntotal}")
    list_vecs = np.empty((ntotal, d), dtype=np.float32)
    list_ids = np.empty(ntotal, dtype=np.int64)
    for i in range(nlist):
        n = int(sizes[i])
        if n:
            a, b = int(offsets[i]), int(offsets[i + 1])
            list_vecs[a:b] = _rd_array(f, "<f4", n * d).reshape(n, d)
            list_ids[a:b] = _rd_array(f, "<i8", n)
    return IVFFlatData(d=d, nlist=nlist, nprobe=nprobe, metric_type=metric, centroids=centroids, offsets=offsets, list_ids=list_ids,
                       list_vecs=list_vecs, is_trained=is_trained)


def write_ivfflat(f: BinaryIO, x: IVFFlatData) -> None:
    """Write what ``faiss.write_index`` writes for this IndexIVFFlat (no direct map, ArrayInvertedLists)."""
    d, nlist = int(x.d), int(x.nlist)
    cen = np.ascontiguousarray(x.centroids, dtype="<f4")
    offsets = np.asarray(x.offsets, dtype=np.int64)
    vecs = np.ascontiguousarray(x.list_vecs, dtype="<f4")
    ids = np.ascontiguousarray(x.list_ids, dtype="<i8")
    if cen.shape != (nlist, d) or offsets.shape != (nlist + 1,) or vecs.shape != (int(offsets[-1]), d) or ids.shape != (int(offsets[-1]),):
        raise FaissFormatError("inconsistent IVFFlatData shapes")
    ntotal = int(offsets[-1])
    f.write(b"IwFl")
    _write_header(f, d, ntotal, x.is_trained, x.metric_type)
    f.write(struct.pack("<QQ", nlist, int(x.nprobe)))
    f.write(b"IxFI" if x.metric_type == METRIC_INNER_PRODUCT else b"IxF2" if x.metric_type == METRIC_L2 else b"IxFl")
    _write_header(f, d, nlist, True, x.metric_type)
    f.write(struct.pack("<Q", nlist * d))
    f.write(cen.tobytes())
    f.write(struct.pack("<BQ", 0, 0))                       # direct map: NoMap, empty array
    f.write(b"ilar")
    f.write(struct.pack("<QQ", nlist, 4 * d))
    sizes = np.diff(offsets).astype("<u8")
    non0 = np.nonzero(sizes)[0]
    if len(non0) > nlist // 2:
        f.write(b"full")
        f.write(struct.pack("<Q", nlist))
        f.write(sizes.tobytes())
    else:
        f.write(b"sprs")
        f.write(struct.pack("<Q", 2 * len(non0)))
        pairs = np.stack([non0.astype("<u8"), sizes[non0]], axis=1)
        f.write(np.ascontiguousarray(pairs, dtype="<u8").tobytes())
    for i in non0:
        a, b = int(offsets[i]), int(offsets[i + 1])
        f.write(vecs[a:b].tobytes())
        f.write(ids[a:b].tobytes())


def sniff(path) -> Optional[bytes]:
    """first 4 bytes of the file (the fourcc), or None when shorter"""
    with open(path, "rb") as f:
        b = f.read(4)
    return b if len(b) == 4 else None
