"""CPU oracle for the IVFFlat / exhaustive inner-product retrieval path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  The arithmetic of this path lives in a third-party dependency that is absent from
``/root/reference`` and from this image: ``faiss-cpu>=1.7.4`` (requirements.txt:2, un-pinned, not
vendored; ``import faiss`` fails here and there is no network).  The reference's own tests hold no
golden neighbours for it (tests/test_models.py:155-246 are shape/ordering/recall properties only).
This file therefore restates the *published* IVFFlat-inner-product algorithm
and is anchored on the reference's call sites and test properties:

* ``normalize_rows``        src/models/faiss_index.py:64-65, 109-110, 141-142  (x / max(‖x‖, 1e-8))
* ``spherical_kmeans``      faiss.IndexIVFFlat.train called at faiss_index.py:73
                            (faiss::Clustering, niter=10, spherical because METRIC_INNER_PRODUCT,
                            max_points_per_centroid=256 sub-sampling, empty-cluster split ±1/1024)
* ``assign`` / ``build_lists``  faiss.IndexIVFFlat.add at faiss_index.py:74 (argmax-IP centroid,
                            vectors appended to their list in insertion order)
* ``ivf_search``            faiss.IndexIVFFlat.search at faiss_index.py:113,145: quantizer top-nprobe
                            by IP, scan those lists in that order, keep the k best by score,
                            descending, id −1 / score −FLT_MAX for unfilled slots
* ``flat_search``           faiss.IndexFlatIP.search semantics (BASELINE cfg 5, exhaustive)
* ``wrapper_search`` / ``wrapper_batch_search``  faiss_index.py:108-123 and :140-153 post-processing

Tie rule: FAISS keeps a k-min-heap and replaces the root only on a strictly greater score, so among
equal scores the earliest scanned wins; the final order among exact ties is heap-dependent and is
excluded from ID equality by BASELINE.json.  Here ties are ordered by scan position (ascending),
which is one of the orders FAISS can produce.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path
from typing import Optional, Tuple

import numpy as np

NEG_SENTINEL = np.float32(-3.4028234663852886e38)  # -FLT_MAX, FAISS heap neutral for IP
_HERE = Path(__file__).resolve().parent


def normalize_rows(x: np.ndarray, eps: float = 1e-8) -> np.ndarray:
    x = np.asarray(x, dtype=np.float32)
    n = np.linalg.norm(x, axis=1, keepdims=True)
    return (x / np.maximum(n, eps)).astype(np.float32)


def assign(x: np.ndarray, centroids: np.ndarray, block: int = 65536) -> np.ndarray:
    """argmax-IP centroid per row (lowest index on ties)."""
    out = np.empty(x.shape[0], dtype=np.int64)
    for s in range(0, x.shape[0], block):
        out[s:s + block] = np.argmax(x[s:s + block] @ centroids.T, axis=1)
    return out


def kmeans_init_indices(n: int, nlist: int, seed: int = 1234) -> np.ndarray:
    """Initial centroids = a seeded random subset of the training points (FAISS: random
    permutation, seed 1234; the exact RNG stream is not reproduced — centroids are an *input* to
    parity runs)."""
    return np.random.default_rng(seed).permutation(n)[:nlist]


def spherical_kmeans(x: np.ndarray, nlist: int, niter: int = 10, seed: int = 1234,
                     max_points_per_centroid: int = 256) -> np.ndarray:
    x = np.asarray(x, dtype=np.float32)
    n = x.shape[0]
    rng = np.random.default_rng(seed + 7)
    if n > max_points_per_centroid * nlist:       # FAISS sub-samples the training set
        x = x[np.random.default_rng(seed + 1).permutation(n)[: max_points_per_centroid * nlist]]
        n = x.shape[0]
    c = x[kmeans_init_indices(n, nlist, seed)].copy()
    for _ in range(niter):
        a = assign(x, c)
        cnt = np.bincount(a, minlength=nlist).astype(np.float32)
        s = np.zeros_like(c)
        np.add.at(s, a, x)
        nz = cnt > 0
        c[nz] = s[nz] / cnt[nz, None]
        # empty-cluster split: copy a populated centroid (prob ∝ size) and perturb ±1/1024
        hs = cnt.copy()
        for ci in np.nonzero(~nz)[0]:
            p = np.maximum(hs - 1, 0)
            cj = int(rng.choice(nlist, p=p / p.sum()))
            c[ci] = c[cj]
            sign = np.where(np.arange(c.shape[1]) % 2 == 0, 1.0, -1.0).astype(np.float32)
            c[ci] *= 1 + sign / 1024
            c[cj] *= 1 - sign / 1024
            hs[ci] = hs[cj] / 2
            hs[cj] -= hs[ci]
        c = normalize_rows(c, 1e-30)              # spherical: renormalise every iteration
    return c.astype(np.float32)


def build_lists(a: np.ndarray, nlist: int) -> Tuple[np.ndarray, np.ndarray]:
    """CSR inverted lists: offsets[nlist+1] and ``order`` = row indices grouped by list, insertion
    (ascending row) order inside each list."""
    order = np.argsort(a, kind="stable").astype(np.int64)
    offsets = np.zeros(nlist + 1, dtype=np.int64)
    np.cumsum(np.bincount(a, minlength=nlist), out=offsets[1:])
    return offsets, order


def coarse_probe(q: np.ndarray, centroids: np.ndarray, nprobe: int) -> np.ndarray:
    """Top-nprobe lists per query by IP, descending, ties → lower list id first."""
    s = q @ centroids.T
    return np.argsort(-s, axis=1, kind="stable")[:, :nprobe]


def coarse_probe_margin(q: np.ndarray, centroids: np.ndarray, nprobe: int) -> np.ndarray:
    """fp64 score gap between the nprobe-th and (nprobe+1)-th centroid per query: queries whose margin is
    below fp32 rounding (~1e-6) may legitimately probe a different list set in another implementation."""
    s = np.sort(q.astype(np.float64) @ centroids.astype(np.float64).T, axis=1)[:, ::-1]
    if nprobe >= s.shape[1]:
        return np.full(s.shape[0], np.inf)
    return s[:, nprobe - 1] - s[:, nprobe]


def assign_margin(x: np.ndarray, centroids: np.ndarray, block: int = 65536) -> np.ndarray:
    """fp64 gap between the best and second-best centroid score per row (∞ with a single centroid)."""
    out = np.empty(x.shape[0])
    if centroids.shape[0] < 2:
        out[:] = np.inf
        return out
    for s0 in range(0, x.shape[0], block):
        sc = np.partition(x[s0:s0 + block].astype(np.float64) @ centroids.astype(np.float64).T, -2, axis=1)
        out[s0:s0 + block] = sc[:, -1] - sc[:, -2]
    return out


def ivf_search(q, centroids, offsets, order, xn, nprobe: int, k: int, dtype=np.float32):
    """Returns (scores[nq,k] f32, idx[nq,k] i64 internal row numbers; −1 / −FLT_MAX padding)."""
    q = np.asarray(q, dtype=np.float32)
    nq = q.shape[0]
    nprobe = min(nprobe, centroids.shape[0])
    probes = coarse_probe(q, centroids, nprobe)
    out_s = np.full((nq, k), NEG_SENTINEL, dtype=np.float32)
    out_i = np.full((nq, k), -1, dtype=np.int64)
    xs = xn.astype(dtype, copy=False)
    for i in range(nq):
        rows = np.concatenate([order[offsets[l]:offsets[l + 1]] for l in probes[i]])
        if rows.size == 0:
            continue
        sc = (xs[rows] @ q[i].astype(dtype)).astype(np.float32)
        top = np.argsort(-sc, kind="stable")[:k]          # score desc, scan position asc
        out_s[i, : top.size] = sc[top]
        out_i[i, : top.size] = rows[top]
    return out_s, out_i


def flat_search(q, xn, k: int, dtype=np.float32, block: int = 262144):
    """Exhaustive inner-product top-k (IndexFlatIP): score desc, row index asc on ties."""
    q = np.asarray(q, dtype=np.float32)
    nq, n = q.shape[0], xn.shape[0]
    k_eff = min(k, n)
    best_s = np.full((nq, 0), 0, dtype=np.float32)
    best_i = np.full((nq, 0), 0, dtype=np.int64)
    for s in range(0, n, block):
        sc = (q.astype(dtype) @ xn[s:s + block].astype(dtype).T).astype(np.float32)
        ids = np.broadcast_to(np.arange(s, s + sc.shape[1], dtype=np.int64), sc.shape)
        cs = np.concatenate([best_s, sc], axis=1)
        ci = np.concatenate([best_i, ids], axis=1)
        top = np.argsort(-cs, axis=1, kind="stable")[:, :k_eff]
        best_s = np.take_along_axis(cs, top, 1)
        best_i = np.take_along_axis(ci, top, 1)
    out_s = np.full((nq, k), NEG_SENTINEL, dtype=np.float32)
    out_i = np.full((nq, k), -1, dtype=np.int64)
    out_s[:, :k_eff] = best_s
    out_i[:, :k_eff] = best_i
    return out_s, out_i


# --- the wrapper's own post-processing (faiss_index.py:108-123, 140-153) -------------------- #
def wrapper_search(scores_row, idx_row, item_ids):
    valid = idx_row >= 0
    return scores_row[valid], item_ids[idx_row[valid]]


def wrapper_batch_search(scores, idx, item_ids):
    mapped = np.where(idx >= 0, item_ids[np.clip(idx, 0, len(item_ids) - 1)], -1)
    return scores, mapped


# --- equivalence modulo ties (BASELINE: "IDs identical except exact-score ties") ------------ #
def assert_topk_equivalent(scores_a, ids_a, scores_b, ids_b, rtol=2e-6, atol=2e-6):
    """a = implementation under test, b = oracle.  Scores must agree slot by slot; ids must agree
    wherever a slot's score is separated from its neighbours (and from the k-th boundary) by more
    than the tolerance; inside a tie group the id *sets* must agree unless the group touches the
    boundary."""
    scores_a, scores_b = np.asarray(scores_a), np.asarray(scores_b)
    ids_a, ids_b = np.asarray(ids_a), np.asarray(ids_b)
    assert scores_a.shape == scores_b.shape and ids_a.shape == ids_b.shape
    np.testing.assert_allclose(scores_a, scores_b, rtol=rtol, atol=atol)
    nq, k = scores_b.shape
    mism = ids_a != ids_b
    if not mism.any():
        return
    for i in np.nonzero(mism.any(1))[0]:
        sb = scores_b[i].astype(np.float64)
        tol = atol + rtol * np.abs(sb)
        # tie groups: consecutive slots whose score gap is within tolerance
        brk = np.nonzero(np.abs(np.diff(sb)) > np.maximum(tol[:-1], tol[1:]))[0] + 1
        for lo, hi in zip(np.r_[0, brk], np.r_[brk, k]):
            if not mism[i, lo:hi].any():
                continue
            assert hi - lo > 1 or hi == k, f"query {i}: id mismatch at untied slot {lo}"
            if hi < k:      # interior tie group: same members, any order
                assert set(ids_a[i, lo:hi].tolist()) == set(ids_b[i, lo:hi].tolist()), \
                    f"query {i}: tie group [{lo},{hi}) has different members"
            # a group touching the k-th boundary may legitimately hold different tied members


# --- C restatement (heap-based, OpenMP over queries, like FAISS) for the CPU baseline ------- #
_LIB = None


def build_c(force: bool = False) -> Path:
    out = _HERE / "_build" / "libivf_oracle.so"
    src = _HERE / "ivf_oracle.c"
    if force or not out.exists() or out.stat().st_mtime < src.stat().st_mtime:
        out.parent.mkdir(exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC",
                               "-o", str(out), str(src), "-lm"])
    return out


def c_lib():
    global _LIB
    if _LIB is None:
        p = _HERE / "_build" / "libivf_oracle.so"
        if not p.exists():
            build_c()
        lib = ctypes.CDLL(str(p))
        P = ctypes.c_void_p
        lib.ivf_oracle_search.argtypes = [P, ctypes.c_int64, ctypes.c_int, P, ctypes.c_int, ctypes.c_int,
                                          P, P, P, ctypes.c_int, P, P, ctypes.c_int]
        lib.ivf_oracle_search.restype = ctypes.c_int
        lib.flat_oracle_search.argtypes = [P, ctypes.c_int64, ctypes.c_int, P, ctypes.c_int64,
                                           ctypes.c_int, P, P, ctypes.c_int]
        lib.flat_oracle_search.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def ivf_search_c(q, centroids, offsets, list_vecs, list_ids, nprobe, k, threads=0):
    """Heap-based search over list-contiguous storage (``list_vecs`` = xn[order], ``list_ids`` =
    order).  Same results as ``ivf_search`` modulo exact ties."""
    q = np.ascontiguousarray(q, np.float32)
    centroids = np.ascontiguousarray(centroids, np.float32)
    offsets = np.ascontiguousarray(offsets, np.int64)
    list_vecs = np.ascontiguousarray(list_vecs, np.float32)
    list_ids = np.ascontiguousarray(list_ids, np.int64)
    nq, D = q.shape
    out_s = np.empty((nq, k), np.float32)
    out_i = np.empty((nq, k), np.int64)
    rc = c_lib().ivf_oracle_search(_p(q), nq, D, _p(centroids), centroids.shape[0], nprobe, _p(offsets),
                                   _p(list_ids), _p(list_vecs), k, _p(out_s), _p(out_i), threads)
    assert rc == 0
    return out_s, out_i


def flat_search_c(q, xn, k, threads=0):
    q = np.ascontiguousarray(q, np.float32)
    xn = np.ascontiguousarray(xn, np.float32)
    nq, D = q.shape
    out_s = np.empty((nq, k), np.float32)
    out_i = np.empty((nq, k), np.int64)
    rc = c_lib().flat_oracle_search(_p(q), nq, D, _p(xn), xn.shape[0], k, _p(out_s), _p(out_i), threads)
    assert rc == 0
    return out_s, out_i
