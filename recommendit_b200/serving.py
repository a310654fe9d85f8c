"""Serving micro-path: one retrieval request = ``model.get_user_embedding(user_id)`` followed by
``faiss_index.search(user_vec, k)`` (src/serving/recommender.py:148-156, 203; src/models/two_tower.py:166-172;
src/models/faiss_index.py:88-124) — as ONE CUDA-graph replay.

At batch 1 the request is bound by launches, small copies and host synchronisations, not by arithmetic: the drop-in classes spend
≈ 10 launches, an H2D of the id, a D2H + H2D of the 256-byte user vector and two host synchronisations (one inside the search
plan, to size the candidate buffer).  Here the user id goes up through a pinned slot, the user tower's output stays on the device
and feeds the search directly, the plan runs in its asynchronous form (candidate buffer sized by the upper bound
nprobe · longest list), and scores + item ids come back through pinned buffers: one graph launch and one synchronisation per
request.  Results are those of the two drop-in calls (same kernels, same order): tests/test_gpu_pipeline.py.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np
import torch

from . import _lib
from ._lib import RB200Error, check, ptr, stream_ptr
from .faiss_index import FAISSIndex
from .two_tower import TwoTowerModel


class UserRecommender:
    def __init__(self, model: TwoTowerModel, index: FAISSIndex, k: int = 500):
        if index.index is None:
            raise RuntimeError("Index not built. Call build_ivf_index() first.")
        self.lib = _lib.load()
        self.model, self.index = model.eval(), index
        self.dev = _lib.require_cuda(model.user_tower.embedding.weight, index.index.centroids)
        self._k_asked = int(k)
        self._id_pinned = torch.zeros(1, dtype=torch.int64).pin_memory()
        self._id_dev = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self._stream = torch.cuda.Stream(self.dev)
        self._graph = None
        self._captured_sig = None
        self._captured_refs = None
        self._size_buffers()

    def _size_buffers(self) -> None:
        """result / plan / candidate buffers for the index as it is now (construction, and again when the index was rebuilt)"""
        st = self.index.index
        self.k = min(self._k_asked, st.ntotal)
        self.nprobe = max(1, min(int(st.nprobe), st.nlist))
        self._q = torch.empty(1, st.d, dtype=torch.float32, device=self.dev)
        self._scores = torch.empty(1, self.k, dtype=torch.float32, device=self.dev)
        self._ids = torch.empty(1, self.k, dtype=torch.int64, device=self.dev)
        self._scores_pinned = torch.empty(1, self.k, dtype=torch.float32).pin_memory()
        self._ids_pinned = torch.empty(1, self.k, dtype=torch.int64).pin_memory()
        self._plan_bytes = self.lib.rb200_ivf_plan_workspace_bytes(1, st.nlist, self.nprobe)
        self._plan = torch.empty(max(self._plan_bytes, 256), dtype=torch.uint8, device=self.dev)
        self._max_cand = self.nprobe * max(st.max_list_len, 1)           # upper bound: no read-back needed
        self._ws_bytes = self.lib.rb200_ivf_search_workspace_bytes(self._max_cand)
        self._ws = torch.empty(max(self._ws_bytes, 256), dtype=torch.uint8, device=self.dev)

    def _signature(self):
        """Everything the captured graph has baked in: the device addresses of the model's parameters and of the index arrays, and the
        sizes the workspaces were carved for.  A rebuilt index (``build_ivf_index`` again), ``model.to()`` or a re-flattened parameter
        block moves them; replaying the old graph would then read freed memory."""
        st = self.index.index
        tensors = [p for p in self.model.user_tower.parameters()] + [st.centroids, st.offsets, st.list_vecs, st.tile_list, st.tile_idx,
                                                                     self.index._list_item_ids]
        return tuple(int(t.data_ptr()) for t in tensors) + (id(st), int(st.max_list_len), int(st.ntotal), int(st.nlist), int(st.d)), tensors

    def _enqueue(self) -> None:
        st = self.index.index
        self._id_dev.copy_(self._id_pinned, non_blocking=True)
        with torch.no_grad():
            u = self.model.user_tower(self._id_dev)                                      # [1, D] on the device
        check(self.lib.rb200_normalize_rows(ptr(u), 1, st.d, 1e-8, ptr(self._q), stream_ptr()), "rb200_normalize_rows")
        check(self.lib.rb200_ivf_search_plan(ptr(self._q), 1, st.d, ptr(st.centroids), st.nlist, self.nprobe, ptr(st.offsets),
                                             ptr(self._plan), self._plan_bytes, None, None, stream_ptr()), "rb200_ivf_search_plan")
        check(self.lib.rb200_ivf_search_run(ptr(self._q), 1, st.d, st.nlist, self.nprobe, ptr(st.offsets), ptr(self.index._list_item_ids),
                                            ptr(st.list_vecs), st.list_vecs.shape[0], st.max_list_len, ptr(st.tile_list), ptr(st.tile_idx),
                                            st.tile_list.numel(), self.k, ptr(self._plan), self._plan_bytes, self._max_cand,
                                            self._max_cand, ptr(self._scores), ptr(self._ids), ptr(self._ws), self._ws_bytes,
                                            stream_ptr()), "rb200_ivf_search_run")
        self._scores_pinned.copy_(self._scores, non_blocking=True)
        self._ids_pinned.copy_(self._ids, non_blocking=True)

    def recommend(self, user_id: int) -> Tuple[np.ndarray, np.ndarray]:
        """→ (scores descending f32[≤k], item ids i64[≤k]) — what ``faiss_index.search(model.get_user_embedding(u), k)`` returns."""
        st = self.index.index
        nprobe = max(1, min(int(st.nprobe), st.nlist))
        if nprobe != self.nprobe:
            raise RB200Error("n_probe changed after the recommender was built: build a new UserRecommender")
        if not 0 <= int(user_id) <= self.model.n_users:          # the captured request cannot raise from the device
            raise IndexError(f"user id {user_id} is outside the embedding table [0, {self.model.n_users}] (torch would raise "
                             "'index out of range in self')")
        self._id_pinned[0] = int(user_id)
        sig, refs = self._signature()
        if self._graph is not None and sig != self._captured_sig:
            self._graph = None                                        # parameters or index arrays moved: size and capture again
            self._size_buffers()
        with torch.cuda.device(self.dev), torch.cuda.stream(self._stream):
            if self._graph is None:
                self._enqueue()                                       # eager once: kernel attributes, allocator
                self._stream.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self._stream):
                    self._enqueue()
                self._graph = g
                self._captured_sig, self._captured_refs = sig, refs   # (strong references: the captured addresses stay valid)
            self._graph.replay()
            self._stream.synchronize()
        s, ids = self._scores_pinned.numpy()[0], self._ids_pinned.numpy()[0]
        valid = ids >= 0
        return s[valid].copy(), ids[valid].copy()


class MicroBatcher:
    """Request micro-batching for the retrieval call of ``recommender.py:304-313`` (second half): concurrent
    callers hand in one query vector each; a worker thread groups whatever arrived within ``max_wait_ms`` (at most ``max_batch``)
    into ONE ``batch_search`` and hands every caller its own row, in the form ``FAISSIndex.search`` returns (padding dropped).

        batcher = MicroBatcher(faiss_index.batch_search, k=500)
        scores, item_ids = batcher.search(user_vec)          # blocks the calling thread only; thread-safe
        batcher.close()

    Pure host logic (``search_fn(np[nq, D], k) -> (np[nq, k], np[nq, k])`` is the only thing it calls), exercised on CPU with a
    stand-in search function in tests/test_serving_host.py.  One request alone waits at most ``max_wait_ms``.
    """

    def __init__(self, search_fn, k: int = 500, max_batch: int = 64, max_wait_ms: float = 0.2):
        import queue
        import threading
        self._search, self.k, self.max_batch, self.max_wait = search_fn, int(k), int(max_batch), float(max_wait_ms) * 1e-3
        self._q: "queue.Queue" = queue.Queue()
        self._closed = False
        self.batches, self.requests = 0, 0
        self._worker = threading.Thread(target=self._run, name="rb200-microbatcher", daemon=True)
        self._worker.start()

    def search(self, query_vector: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        import concurrent.futures
        if self._closed:
            raise RuntimeError("MicroBatcher is closed")
        fut: "concurrent.futures.Future" = concurrent.futures.Future()
        self._q.put((np.asarray(query_vector, dtype=np.float32).reshape(-1), fut))
        return fut.result()

    def _run(self) -> None:
        import queue
        import time
        while True:
            item = self._q.get()
            if item is None:
                return
            group = [item]
            deadline = time.perf_counter() + self.max_wait
            while len(group) < self.max_batch:
                left = deadline - time.perf_counter()
                if left <= 0:
                    break
                try:
                    nxt = self._q.get(timeout=left)
                except queue.Empty:
                    break
                if nxt is None:
                    self._q.put(None)                      # let the outer loop see the shutdown after this group
                    break
                group.append(nxt)
            try:
                scores, ids = self._search(np.stack([g[0] for g in group]), self.k)
                self.batches += 1
                self.requests += len(group)
                for row, (_, fut) in enumerate(group):
                    valid = ids[row] >= 0
                    fut.set_result((scores[row][valid].copy(), ids[row][valid].copy()))
            except Exception as e:                         # a failed batch fails its callers, not the worker
                for _, fut in group:
                    if not fut.done():
                        fut.set_exception(e)

    def close(self) -> None:
        if not self._closed:
            self._closed = True
            self._q.put(None)
            self._worker.join(timeout=5)
