"""B200 drop-in for the reference's ``src/models/faiss_index.py`` — no FAISS involved.

``FAISSIndex`` keeps the reference's name and call surface (``build_ivf_index`` / ``search`` /
``batch_search`` / ``save`` / ``load`` / ``stats`` / ``set_n_probe``; attributes ``index.ntotal``,
``index.nprobe``, ``item_ids``, ``embed_dim``, ``n_lists``, ``n_probe``) so ``build_index.py:128-138``,
``serving/recommender.py:155-156,311-313`` and ``run_pipeline.py:146-178`` use it unchanged.  What FAISS's
``IndexIVFFlat(IndexFlatIP, METRIC_INNER_PRODUCT)`` did on the CPU is done by ``librb200.so``:

=============================  ==========================================================
reference call                 replaced by
=============================  ==========================================================
``index.train`` (:73)          spherical k-means: ``rb200_ivf_assign`` + ``rb200_kmeans_update``
``index.add`` (:74)            ``rb200_ivf_assign`` + ``rb200_ivf_build`` (CSR inverted lists in HBM)
``index.search`` (:113,:145)   ``rb200_ivf_search_plan`` + ``rb200_ivf_search_run``
numpy renormalisation          ``rb200_normalize_rows``
=============================  ==========================================================

The database, the centroids and the inverted lists stay resident in HBM; queries come in and results
go out as host NumPy arrays exactly as in the reference.
"""
from __future__ import annotations

import ctypes as C
import logging
import pickle
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import RB200Error, check, ptr, stream_ptr, workspace

logger = logging.getLogger(__name__)

KMEANS_ITERS = 10          # faiss::ClusteringParameters default niter for the IVF coarse quantizer
KMEANS_SEED = 1234         # faiss default seed
MAX_POINTS_PER_CENTROID = 256
_MAGIC = b"RB200IVF1"


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise RB200Error("FAISSIndex (recommendit_b200) needs a CUDA device: the IVF kernels are sm_100a only, "
                         "there is no CPU fallback.")
    return torch.device("cuda", torch.cuda.current_device())


class _IVFState:
    """What ``FAISSIndex.index`` points at: device-resident IVFFlat state with the two attributes the
    reference touches on a faiss index (``ntotal``, ``nprobe``)."""

    def __init__(self, d: int, nlist: int, nprobe: int):
        self.d = d
        self.nlist = nlist
        self.nprobe = nprobe
        self.ntotal = 0
        self.centroids: Optional[torch.Tensor] = None   # [nlist, d] f32
        self.offsets: Optional[torch.Tensor] = None     # [nlist+1] i64
        self.list_ids: Optional[torch.Tensor] = None    # [ntotal] i64 internal row numbers, list-contiguous
        self.list_vecs: Optional[torch.Tensor] = None   # [ntotal, d] f32, list-contiguous
        self.max_list_len = 0
        self.is_trained = False

    # ---- kernels -------------------------------------------------------------------------- #
    def assign(self, x: torch.Tensor) -> torch.Tensor:
        lib = _lib.load()
        out = torch.empty(x.shape[0], dtype=torch.int32, device=x.device)
        check(lib.rb200_ivf_assign(ptr(x), x.shape[0], self.d, ptr(self.centroids), self.nlist, ptr(out), None,
                                   stream_ptr()), "rb200_ivf_assign")
        return out

    def train(self, x: torch.Tensor, niter: int = KMEANS_ITERS, seed: int = KMEANS_SEED) -> None:
        """Spherical k-means (metric = inner product ⇒ centroids renormalised every iteration)."""
        lib = _lib.load()
        n = x.shape[0]
        if n > MAX_POINTS_PER_CENTROID * self.nlist:     # FAISS sub-samples large training sets
            sel = np.random.default_rng(seed + 1).permutation(n)[: MAX_POINTS_PER_CENTROID * self.nlist]
            x = x[torch.as_tensor(sel, device=x.device)].contiguous()
            n = x.shape[0]
        init = np.random.default_rng(seed).permutation(n)[: self.nlist]
        if len(init) < self.nlist:                       # fewer points than lists: repeat (degenerate case)
            init = np.resize(init, self.nlist)
        self.centroids = x[torch.as_tensor(init, device=x.device)].contiguous().clone()
        counts = torch.empty(self.nlist, dtype=torch.int32, device=x.device)
        wsb = lib.rb200_kmeans_update_workspace_bytes(n, self.nlist)
        ws = workspace(wsb, x.device)
        rng = np.random.default_rng(seed + 7)
        for _ in range(niter):
            a = self.assign(x)
            check(lib.rb200_kmeans_update(ptr(x), n, self.d, ptr(a), self.nlist, ptr(self.centroids), ptr(counts),
                                          ptr(ws), wsb, stream_ptr()), "rb200_kmeans_update")
            cnt = counts.cpu().numpy().astype(np.float64)
            empty = np.nonzero(cnt == 0)[0]
            if len(empty):                               # FAISS split_clusters: clone a big list, perturb ±1/1024
                cen = self.centroids.cpu().numpy()
                sign = np.where(np.arange(self.d) % 2 == 0, 1.0, -1.0).astype(np.float32)
                for ci in empty:
                    p = np.maximum(cnt - 1, 0)
                    if p.sum() <= 0:
                        break
                    cj = int(rng.choice(self.nlist, p=p / p.sum()))
                    cen[ci] = cen[cj] * (1 + sign / 1024)
                    cen[cj] = cen[cj] * (1 - sign / 1024)
                    cnt[ci] = cnt[cj] / 2
                    cnt[cj] -= cnt[ci]
                self.centroids = torch.as_tensor(cen, device=x.device)
                check(lib.rb200_normalize_rows(ptr(self.centroids), self.nlist, self.d, 1e-30, ptr(self.centroids),
                                               stream_ptr()), "rb200_normalize_rows")
        self.is_trained = True

    def add(self, x: torch.Tensor) -> None:
        lib = _lib.load()
        n = x.shape[0]
        a = self.assign(x)
        self.offsets = torch.empty(self.nlist + 1, dtype=torch.int64, device=x.device)
        self.list_ids = torch.empty(n, dtype=torch.int64, device=x.device)
        self.list_vecs = torch.empty(n, self.d, dtype=torch.float32, device=x.device)
        wsb = lib.rb200_ivf_build_workspace_bytes(n, self.nlist)
        ws = workspace(wsb, x.device)
        check(lib.rb200_ivf_build(ptr(x), n, self.d, ptr(a), self.nlist, ptr(self.offsets), ptr(self.list_ids),
                                  ptr(self.list_vecs), ptr(ws), wsb, stream_ptr()), "rb200_ivf_build")
        self.ntotal = n
        self.finalize()

    def finalize(self) -> None:
        """Derived search-side tables (depend only on ``offsets``): longest list and the (list, 64-vector tile) work items
        of the list scan."""
        lens = self.offsets[1:] - self.offsets[:-1]
        self.max_list_len = int(lens.max().item()) if lens.numel() else 0
        n_tile = (lens + 63) // 64
        first = torch.cumsum(n_tile, 0) - n_tile
        lists = torch.arange(self.nlist, device=lens.device)
        self.tile_list = torch.repeat_interleave(lists, n_tile).to(torch.int32).contiguous()
        self.tile_idx = (torch.arange(int(n_tile.sum().item()), device=lens.device) -
                         torch.repeat_interleave(first, n_tile)).to(torch.int32).contiguous()

    def search_device(self, q: torch.Tensor, k: int, id_table: Optional[torch.Tensor] = None
                      ) -> Tuple[torch.Tensor, torch.Tensor]:
        """q: normalised [nq, d] f32 on the device → (scores [nq,k] f32, rows [nq,k] i64), -FLT_MAX / -1 padded.
        ``id_table`` (list-contiguous, same order as ``list_ids``) makes the kernel emit those ids instead of
        internal row numbers."""
        lib = _lib.load()
        ids_src = self.list_ids if id_table is None else id_table
        nq = q.shape[0]
        nprobe = max(1, min(int(self.nprobe), self.nlist))
        pb = lib.rb200_ivf_plan_workspace_bytes(nq, self.nlist, nprobe)
        plan = workspace(pb, q.device)
        total, mx = C.c_int64(0), C.c_int64(0)
        check(lib.rb200_ivf_search_plan(ptr(q), nq, self.d, ptr(self.centroids), self.nlist, nprobe, ptr(self.offsets),
                                        ptr(plan), pb, C.byref(total), C.byref(mx), stream_ptr()), "rb200_ivf_search_plan")
        wb = lib.rb200_ivf_search_workspace_bytes(total.value)
        ws = workspace(wb, q.device)
        scores = torch.empty(nq, k, dtype=torch.float32, device=q.device)
        rows = torch.empty(nq, k, dtype=torch.int64, device=q.device)
        check(lib.rb200_ivf_search_run(ptr(q), nq, self.d, self.nlist, nprobe, ptr(self.offsets), ptr(ids_src),
                                       ptr(self.list_vecs), self.list_vecs.shape[0], self.max_list_len, ptr(self.tile_list), ptr(self.tile_idx),
                                       self.tile_list.numel(), k, ptr(plan), pb, total.value, mx.value,
                                       ptr(scores), ptr(rows), ptr(ws), wb, stream_ptr()), "rb200_ivf_search_run")
        self._last_plan = (plan, pb, nq, nprobe)          # for check_last_search()
        return scores, rows

    def check_last_search(self) -> None:
        """SYNCHRONISES the stream and raises if a tensor-core pipeline of the last ``search_device`` timed out (garbage must not
        pass as a result).  The host-facing searches call it where they wait for their device→host copies anyway."""
        last = getattr(self, "_last_plan", None)
        if last is None:
            return
        plan, pb, nq, nprobe = last
        check(_lib.load().rb200_ivf_search_status(ptr(plan), pb, nq, self.nlist, nprobe, stream_ptr()), "rb200_ivf_search_status")

    def search(self, queries: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
        """faiss-style ``index.search(x, k)`` on host arrays (already normalised)."""
        q = torch.as_tensor(np.ascontiguousarray(queries, dtype=np.float32), device=self.centroids.device)
        s, r = self.search_device(q, k)
        self.check_last_search()
        return s.cpu().numpy(), r.cpu().numpy()


def _normalize_device(x: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    lib = _lib.load()
    out = torch.empty_like(x)
    check(lib.rb200_normalize_rows(ptr(x), x.shape[0], x.shape[1], eps, ptr(out), stream_ptr()), "rb200_normalize_rows")
    return out


class FAISSIndex:
    """IVFFlat inner-product index over L2-normalised vectors; scores are cosine similarities."""

    def __init__(self, embed_dim: int = 64, n_lists: int = 100, n_probe: int = 10):
        self.embed_dim = embed_dim
        self.n_lists = n_lists
        self.n_probe = n_probe
        self.index: Optional[_IVFState] = None
        self.item_ids: Optional[np.ndarray] = None            # internal row → item id
        self._item_id_to_faiss_idx: Dict[int, int] = {}

    # ---- construction (faiss_index.py:45-82) ----------------------------------------------- #
    def build_ivf_index(self, embeddings: np.ndarray, item_ids: List[int], centroids: Optional[np.ndarray] = None) -> None:
        """``centroids`` (optional, [n_lists, embed_dim]) skips k-means and uses the given coarse quantizer —
        parity runs feed the same centroids to this index and to the CPU oracle."""
        assert embeddings.dtype == np.float32, "Embeddings must be float32"
        assert embeddings.shape[1] == self.embed_dim, f"Expected embed_dim={self.embed_dim}, got {embeddings.shape[1]}"
        dev = _device()
        n = embeddings.shape[0]
        with torch.cuda.device(dev):
            x = _normalize_device(torch.as_tensor(np.ascontiguousarray(embeddings), device=dev))   # :64-65
            st = _IVFState(self.embed_dim, self.n_lists, self.n_probe)
            if centroids is not None:
                assert centroids.shape == (self.n_lists, self.embed_dim)
                st.centroids = torch.as_tensor(np.ascontiguousarray(centroids, dtype=np.float32), device=dev)
                st.is_trained = True
            else:
                logger.info("Training IVF coarse quantizer on %d vectors (n_lists=%d)...", n, self.n_lists)
                st.train(x)
            st.add(x)
        self.index = st
        self.item_ids = np.array(item_ids, dtype=np.int64)
        self._item_id_to_faiss_idx = {int(iid): idx for idx, iid in enumerate(item_ids)}
        self._bind_item_ids()
        logger.info("IVF index built: %d vectors, %d lists, probe=%d", st.ntotal, self.n_lists, self.n_probe)

    def _bind_item_ids(self) -> None:
        """internal row → item id (faiss_index.py:76, :118-123, :147-152) folded into the index: the list entries carry the
        item id, so the top-k kernel emits catalog ids directly and no per-search host mapping is needed."""
        st = self.index
        ids_dev = torch.as_tensor(np.asarray(self.item_ids, dtype=np.int64), device=st.list_ids.device)
        self._list_item_ids = ids_dev[st.list_ids].contiguous()

    # ---- search (faiss_index.py:88-153) ------------------------------------------------------ #
    def _search_normalised(self, queries: np.ndarray, k: int):
        """host queries → (scores, item ids) as host arrays; -FLT_MAX / -1 padded.  Results come back through pinned
        buffers from torch's pinned caching allocator (one async D2H each)."""
        st = self.index
        dev = st.centroids.device
        with torch.cuda.device(dev):
            q = _normalize_device(torch.as_tensor(np.ascontiguousarray(queries, dtype=np.float32), device=dev))
            s, r = st.search_device(q, k, id_table=self._list_item_ids)
            hs = torch.empty(s.shape, dtype=s.dtype, pin_memory=True)
            hr = torch.empty(r.shape, dtype=r.dtype, pin_memory=True)
            hs.copy_(s, non_blocking=True)
            hr.copy_(r, non_blocking=True)
            st.check_last_search()                       # (synchronises the stream: the copies have landed)
            return hs.numpy(), hr.numpy()

    def search(self, query_vector: np.ndarray, k: int = 500) -> Tuple[np.ndarray, np.ndarray]:
        if self.index is None:
            raise RuntimeError("Index not built. Call build_ivf_index() first.")
        query = np.atleast_2d(query_vector).astype(np.float32)
        k = min(k, self.index.ntotal)
        distances, ids = self._search_normalised(query[:1], k)      # like the reference, only row 0 is used (:115-116)
        distances, ids = distances[0], ids[0]
        valid = ids >= 0 if (self.item_ids >= 0).all() else distances > -3e38
        return distances[valid].copy(), ids[valid].copy()

    def batch_search(self, query_vectors: np.ndarray, k: int = 500) -> Tuple[np.ndarray, np.ndarray]:
        if self.index is None:
            raise RuntimeError("Index not built.")
        k = min(k, self.index.ntotal)
        return self._search_normalised(query_vectors, k)                # padding slots already carry id -1

    # ---- persistence (faiss_index.py:159-205) ---------------------------------------------- #
    def save(self, path: str, file_format: str = "faiss") -> None:
        """``path`` holds the index in FAISS's own on-disk format — what ``faiss.write_index(IndexIVFFlat)`` writes at
        faiss_index.py:164 ("IwFl": flat IP quantizer + array inverted lists, ``faiss_io.py``) — so index files move between the
        reference and this package in both directions; ``path.with_suffix('.meta.pkl')`` is the same pickle sidecar the
        reference writes.  ``file_format='rb200'`` writes round 1's private container (still readable by ``load``)."""
        if self.index is None:
            raise RuntimeError("Index not built.")
        if file_format not in ("faiss", "rb200"):
            raise ValueError("file_format must be 'faiss' or 'rb200'")
        save_path = Path(path)
        save_path.parent.mkdir(parents=True, exist_ok=True)
        st = self.index
        with open(save_path, "wb") as f:
            if file_format == "faiss":
                from . import faiss_io
                faiss_io.write_ivfflat(f, faiss_io.IVFFlatData(
                    d=st.d, nlist=st.nlist, nprobe=max(1, int(st.nprobe)), metric_type=faiss_io.METRIC_INNER_PRODUCT,
                    centroids=st.centroids.cpu().numpy(), offsets=st.offsets.cpu().numpy(), list_ids=st.list_ids.cpu().numpy(),
                    list_vecs=st.list_vecs.cpu().numpy(), is_trained=True))
            else:
                f.write(_MAGIC)
                np.savez(f, centroids=st.centroids.cpu().numpy(), offsets=st.offsets.cpu().numpy(),
                         list_ids=st.list_ids.cpu().numpy(), list_vecs=st.list_vecs.cpu().numpy(),
                         dims=np.array([st.d, st.nlist, st.ntotal, st.max_list_len], dtype=np.int64))
        with open(save_path.with_suffix(".meta.pkl"), "wb") as f:
            pickle.dump({"item_ids": self.item_ids, "item_id_to_faiss_idx": self._item_id_to_faiss_idx,
                         "embed_dim": self.embed_dim, "n_lists": self.n_lists, "n_probe": self.n_probe}, f)
        logger.info("Saved IVF index to %s", save_path)

    @classmethod
    def load(cls, path: str) -> "FAISSIndex":
        """Reads an index file written by this class OR by the reference (``faiss.write_index`` of its IndexIVFFlat) plus the
        ``.meta.pkl`` sidecar both write (faiss_index.py:184-197)."""
        load_path = Path(path)
        if not load_path.exists():
            raise FileNotFoundError(f"FAISS index not found at {load_path}")
        with open(load_path.with_suffix(".meta.pkl"), "rb") as f:
            meta = pickle.load(f)
        obj = cls(embed_dim=meta["embed_dim"], n_lists=meta["n_lists"], n_probe=meta["n_probe"])
        dev = _device()
        from . import faiss_io
        with open(load_path, "rb") as f:
            head = f.read(len(_MAGIC))
            f.seek(0)
            if head == _MAGIC:
                f.read(len(_MAGIC))
                z = np.load(f)
                d, nlist, ntotal, mx = (int(v) for v in z["dims"])
                cen, off, lid, lvec = z["centroids"], z["offsets"], z["list_ids"], z["list_vecs"]
            elif head[:4] == b"IwFl":
                x = faiss_io.read_ivfflat(f)
                if x.metric_type != faiss_io.METRIC_INNER_PRODUCT:
                    raise RB200Error(f"{load_path}: the index uses metric {x.metric_type}; this class serves inner-product indexes")
                d, nlist, ntotal = x.d, x.nlist, x.ntotal
                cen, off, lid, lvec = x.centroids, x.offsets, x.list_ids, x.list_vecs
            else:
                raise RB200Error(f"{load_path} is neither a FAISS IndexIVFFlat file ('IwFl') nor a recommendit_b200 IVF index file")
        st = _IVFState(d, nlist, meta["n_probe"])                          # nprobe from the sidecar, as faiss_index.py:197
        st.centroids = torch.as_tensor(np.ascontiguousarray(cen, dtype=np.float32), device=dev)
        st.offsets = torch.as_tensor(np.ascontiguousarray(off, dtype=np.int64), device=dev)
        st.list_ids = torch.as_tensor(np.ascontiguousarray(lid, dtype=np.int64), device=dev)
        st.list_vecs = torch.as_tensor(np.ascontiguousarray(lvec, dtype=np.float32), device=dev)
        st.ntotal, st.is_trained = ntotal, True
        st.finalize()
        obj.index = st
        obj.item_ids = meta["item_ids"]
        obj._item_id_to_faiss_idx = meta["item_id_to_faiss_idx"]
        obj._bind_item_ids()
        logger.info("Loaded IVF index from %s: %d vectors, dim=%d", load_path, st.ntotal, obj.embed_dim)
        return obj

    # ---- utilities (faiss_index.py:211-228) --------------------------------------------------- #
    def stats(self) -> Dict:
        if self.index is None:
            return {"status": "not built"}
        return {"n_vectors": int(self.index.ntotal), "embed_dim": self.embed_dim, "n_lists": self.n_lists,
                "n_probe": self.n_probe, "metric": "inner_product",
                "n_item_ids": len(self.item_ids) if self.item_ids is not None else 0}

    def set_n_probe(self, n_probe: int) -> None:
        self.n_probe = n_probe
        if self.index is not None:
            self.index.nprobe = n_probe


def flat_search(queries: torch.Tensor, database: torch.Tensor, k: int, id_base: int = 0
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Exhaustive inner-product top-k on device tensors (``faiss.IndexFlatIP.search`` semantics; BASELINE cfg 5).
    Returns (scores [nq,k], ids [nq,k] = id_base + row), -FLT_MAX / -1 padded when k > rows."""
    lib = _lib.load()
    _lib.require_cuda(queries, database)
    nq, D = queries.shape
    n = database.shape[0]
    wb = lib.rb200_flat_search_workspace_bytes(nq, n, k)
    ws = workspace(wb, queries.device)
    scores = torch.empty(nq, k, dtype=torch.float32, device=queries.device)
    ids = torch.empty(nq, k, dtype=torch.int64, device=queries.device)
    with torch.cuda.device(queries.device):
        check(lib.rb200_flat_search(ptr(queries), nq, ptr(database), n, D, k, id_base, ptr(scores), ptr(ids), ptr(ws), wb,
                                    stream_ptr()), "rb200_flat_search")
    return scores, ids


def topk_merge(scores: torch.Tensor, ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge per-shard sorted top-k lists: scores/ids are [parts, nq, k] → ([nq,k], [nq,k])."""
    lib = _lib.load()
    _lib.require_cuda(scores, ids)
    parts, nq, k = scores.shape
    scores, ids = scores.contiguous(), ids.contiguous()
    wb = lib.rb200_topk_merge_workspace_bytes(parts, nq, k)
    ws = workspace(wb, scores.device)
    out_s = torch.empty(nq, k, dtype=torch.float32, device=scores.device)
    out_i = torch.empty(nq, k, dtype=torch.int64, device=scores.device)
    with torch.cuda.device(scores.device):
        check(lib.rb200_topk_merge(ptr(scores), ptr(ids), parts, nq, k, ptr(out_s), ptr(out_i), ptr(ws), wb, stream_ptr()),
              "rb200_topk_merge")
    return out_s, out_i


def scores_nt(a: torch.Tensor, b: torch.Tensor, mode: int = 2) -> torch.Tensor:
    """``a @ b.T`` on the tcgen05 tensor cores (``rb200_gemm_nt``): mode 1 = TF32, mode 2 = 3xTF32 (fp32-grade)."""
    lib = _lib.load()
    _lib.require_cuda(a, b)
    a, b = a.contiguous(), b.contiguous()
    M, K = a.shape
    N = b.shape[0]
    out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    err = torch.zeros(1, dtype=torch.int32, device=a.device)
    with torch.cuda.device(a.device):
        check(lib.rb200_gemm_nt(ptr(a), M, ptr(b), N, K, mode, ptr(out), N, ptr(err), stream_ptr()), "rb200_gemm_nt")
    if int(err.item()) != 0:
        raise RB200Error("rb200_gemm_nt: tensor-core pipeline timed out (error flag %d)" % int(err.item()))
    return out

