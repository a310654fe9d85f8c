"""Fused BPR training step: the body of the reference's hot loop (``src/training/train_embeddings.py:183-192``)
as ONE call into ``librb200.so`` (``rb200_bpr_step``), replayed as a CUDA graph.

    trainer = FusedBPRTrainer(model, lr=1e-3, weight_decay=1e-5)        # == Adam(...) + clip_grad_norm_(1.0)
    loss = trainer.step_host(user_ids, pos_ids, pos_genres, neg_ids, neg_genres)   # host (pinned) batch in, float out

Per step the graph runs: 3 tower evaluations (1 launch) → pairwise ``bpr_loss`` (or ``in_batch_bpr_loss``) with its
gradient → tower backward → deterministic sorted-segment scatter of the row gradients → global-norm clip → Adam.
``adam_mode='dense'`` reproduces the reference trajectory exactly (every table row moves each step because of the
coupled weight decay); ``adam_mode='rows'`` updates only the rows the batch touched.

The unchanged reference caller cannot reach this entry point (it owns its own ``torch.optim.Adam``); it goes through
the autograd path in ``two_tower.py``.  This class is the alternative trainer for users who can change two lines.
"""
from __future__ import annotations

import ctypes as C
import os
import math
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from ._lib import OptState, RB200Error, Sampler, StepParams, StepViews, check, ptr, require_cuda, stream_ptr
from .two_tower import TwoTowerModel, inbatch_mode_for


def _flat_len(tower) -> int:
    return sum(p.numel() for p in tower.mlp.parameters())


class FusedBPRTrainer:
    def __init__(self, model: TwoTowerModel, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-5, max_norm: float = 1.0, adam_mode: str = "dense", loss: str = "bpr",
                 inbatch_mode: Optional[int] = None, item_extra_table: Optional[torch.Tensor] = None,
                 use_cuda_graph: bool = True, seed: Optional[int] = None, tower_mode=None):
        self.lib = _lib.load()
        self.model = model
        self.dev = require_cuda(*list(model.parameters()))
        if adam_mode not in ("dense", "rows"):
            raise ValueError("adam_mode must be 'dense' (reference-exact) or 'rows' (touched rows only)")
        if loss not in ("bpr", "in_batch"):
            raise ValueError("loss must be 'bpr' or 'in_batch'")
        self.adam_mode, self.loss_kind = adam_mode, (0 if loss == "bpr" else 1)
        self.inbatch_mode = inbatch_mode          # None → auto (tcgen05 3xTF32 where built), resolved per batch size
        self.use_graph = use_cuda_graph
        from .two_tower import tower_mode_for
        self.tower_mode = tower_mode_for(model.embed_dim, model.user_tower.mlp[0].out_features, model.item_tower.extra_dim, tower_mode)
        self.seed = (torch.initial_seed() if seed is None else seed) & 0x7FFFFFFFFFFFFFFF
        self.D, self.H = model.embed_dim, model.user_tower.mlp[0].out_features
        self.E = model.item_tower.extra_dim
        self.item_extra_table = None
        if item_extra_table is not None:
            t = item_extra_table.to(self.dev, torch.float32).contiguous()
            if t.shape != (model.n_items + 1, self.E):
                raise ValueError(f"item_extra_table must be [{model.n_items + 1}, {self.E}]")
            self.item_extra_table = t
        self._flatten()
        f32 = dict(dtype=torch.float32, device=self.dev)
        ut, it = model.user_tower.embedding.weight, model.item_tower.embedding.weight
        self.state = {
            "user_table_m": torch.zeros_like(ut), "user_table_v": torch.zeros_like(ut),
            "item_table_m": torch.zeros_like(it), "item_table_v": torch.zeros_like(it),
            "user_mlp_m": torch.zeros(self._user_flat.numel(), **f32), "user_mlp_v": torch.zeros(self._user_flat.numel(), **f32),
            "item_mlp_m": torch.zeros(self._item_flat.numel(), **f32), "item_mlp_v": torch.zeros(self._item_flat.numel(), **f32),
        }
        self.user_slot = torch.full((ut.shape[0],), -1, dtype=torch.int32, device=self.dev)
        self.item_slot = torch.full((it.shape[0],), -1, dtype=torch.int32, device=self.dev)
        self.opt_dev = torch.zeros(128, dtype=torch.uint8, device=self.dev)
        self._opt_host = OptState()
        h = self._opt_host
        h.lr, h.beta1, h.beta2, h.eps, h.weight_decay, h.max_norm = lr, betas[0], betas[1], eps, weight_decay, max_norm
        h.one_minus_beta1, h.one_minus_beta2, h.beta2_f = 1 - betas[0], 1 - betas[1], betas[1]
        h.clip_coef, h.step = 1.0, 0
        self._push_opt()
        self.loss_dev = torch.zeros(1, **f32)
        self.err_flag = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self._B = None
        self._graph = None
        self._steps_done = 0
        self._sampler = None          # (Sampler struct, producer) when the step produces its own next batch

    def attach_producer(self, sampler: Optional[Sampler], owner=None) -> None:
        """Let the step sample its own NEXT batch on the device (``rb200_step_params.next_batch``, csrc/sampler.cu): the
        ids never come from the host and an epoch is nothing but graph replays.  ``None`` detaches."""
        if sampler is not None and self.loss_kind != 0:
            raise RB200Error("the device-side producer yields (user, positive, negative) triples: loss must be 'bpr'")
        if sampler is not None and self.E and self.item_extra_table is None:
            raise RB200Error("the device-side producer needs a trainer built with item_extra_table (genres by item id)")
        self._sampler = None if sampler is None else (sampler, owner)
        self._graph = None

    # ---- parameter plumbing ------------------------------------------------------------------- #
    def _flatten(self) -> None:
        """Make the 4 MLP tensors of each tower views of ONE flat block [W1 | b1 | W2 | b2] (what the step kernel and
        the fused Adam walk); the nn.Parameters keep working because only ``.data`` is re-pointed."""
        flats = []
        for tower in (self.model.user_tower, self.model.item_tower):
            ps = [tower.mlp[0].weight, tower.mlp[0].bias, tower.mlp[3].weight, tower.mlp[3].bias]
            flat = torch.cat([p.detach().reshape(-1).float() for p in ps]).contiguous()
            o = 0
            for p in ps:
                p.data = flat[o:o + p.numel()].view(p.shape)
                o += p.numel()
            flats.append(flat)
        self._user_flat, self._item_flat = flats
        for emb in (self.model.user_tower.embedding, self.model.item_tower.embedding):
            if not emb.weight.is_contiguous() or emb.weight.dtype != torch.float32:
                emb.weight.data = emb.weight.data.float().contiguous()

    def _is_flat(self) -> bool:
        return (self.model.user_tower.mlp[0].weight.data_ptr() == self._user_flat.data_ptr()
                and self.model.item_tower.mlp[0].weight.data_ptr() == self._item_flat.data_ptr())

    def _push_opt(self) -> None:
        host = torch.frombuffer(bytearray(bytes(self._opt_host)), dtype=torch.uint8)
        self.opt_dev.copy_(host)

    def _pull_opt(self) -> OptState:
        raw = bytes(self.opt_dev.cpu().numpy().tobytes())
        return OptState.from_buffer_copy(raw)

    def set_lr(self, lr: float) -> None:
        """For ``CosineAnnealingLR``-style schedules (the reference steps it once per epoch)."""
        cur = self._pull_opt()
        cur.lr = lr
        self._opt_host = cur
        self._push_opt()

    @property
    def opt_state(self) -> OptState:
        return self._pull_opt()

    # ---- batch buffers ---------------------------------------------------------------------------- #
    def _alloc(self, B: int) -> None:
        self._B = B
        E = 0 if self.item_extra_table is not None else self.E
        n_i64 = 3 * B
        self.batch_bytes = n_i64 * 8 + 2 * B * E * 4
        # one packed staging buffer: [user_ids | pos_ids | neg_ids | pos_extra | neg_extra]
        self.batch_dev = torch.zeros(self.batch_bytes, dtype=torch.uint8, device=self.dev)
        self.batch_pinned = torch.zeros(self.batch_bytes, dtype=torch.uint8).pin_memory()
        ids = self.batch_dev[: n_i64 * 8].view(torch.int64)
        self.user_ids, self.pos_ids, self.neg_ids = ids[:B], ids[B:2 * B], ids[2 * B:]
        if E:
            ex = self.batch_dev[n_i64 * 8:].view(torch.float32)
            self.pos_extra, self.neg_extra = ex[: B * E].view(B, E), ex[B * E:].view(B, E)
        else:
            self.pos_extra = self.neg_extra = self.item_extra_table
        wsb = self.lib.rb200_bpr_step_workspace_bytes(B, self.D, self.H, self.E, self.model.n_users + 1,
                                                      self.model.n_items + 1, self.loss_kind)
        self.ws = torch.empty(wsb, dtype=torch.uint8, device=self.dev)
        self._graph = None
        self._params = None

    def pack_host(self, user_ids, pos_ids, pos_genres, neg_ids, neg_genres, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Pack one batch (NumPy arrays or CPU tensors) into the staging layout; returns a pinned uint8 tensor."""
        B = len(user_ids)
        if self._B != B:
            self._alloc(B)
        buf = self.batch_pinned if out is None else out
        nb = buf.numpy()
        ids = nb[: 3 * B * 8].view(np.int64)
        ids[:B], ids[B:2 * B], ids[2 * B:] = np.asarray(user_ids), np.asarray(pos_ids), np.asarray(neg_ids)
        if self.item_extra_table is None and self.E:
            ex = nb[3 * B * 8:].view(np.float32)
            ex[: B * self.E] = np.asarray(pos_genres, dtype=np.float32).reshape(-1)
            ex[B * self.E:] = np.asarray(neg_genres, dtype=np.float32).reshape(-1)
        return buf

    # ---- the step ------------------------------------------------------------------------------------ #
    def _make_params(self) -> StepParams:
        m, s = self.model, self.state
        p = StepParams()
        p.D, p.H, p.extra_dim, p.B = self.D, self.H, self.E, self._B
        p.n_user_rows, p.n_item_rows = m.n_users + 1, m.n_items + 1
        p.user_table, p.user_table_m, p.user_table_v = ptr(m.user_tower.embedding.weight), ptr(s["user_table_m"]), ptr(s["user_table_v"])
        p.item_table, p.item_table_m, p.item_table_v = ptr(m.item_tower.embedding.weight), ptr(s["item_table_m"]), ptr(s["item_table_v"])
        p.user_mlp, p.user_mlp_m, p.user_mlp_v = ptr(self._user_flat), ptr(s["user_mlp_m"]), ptr(s["user_mlp_v"])
        p.item_mlp, p.item_mlp_m, p.item_mlp_v = ptr(self._item_flat), ptr(s["item_mlp_m"]), ptr(s["item_mlp_v"])
        p.user_row_slot, p.item_row_slot, p.opt = ptr(self.user_slot), ptr(self.item_slot), ptr(self.opt_dev)
        p.user_ids, p.pos_ids, p.neg_ids = ptr(self.user_ids), ptr(self.pos_ids), ptr(self.neg_ids)
        p.pos_extra, p.neg_extra = ptr(self.pos_extra), ptr(self.neg_extra)
        p.extra_by_id = 1 if self.item_extra_table is not None else 0
        p.loss_kind, p.inbatch_mode = self.loss_kind, inbatch_mode_for(self._B, self.D, self.inbatch_mode)
        p.tower_mode = self.tower_mode
        p.adam_mode = 0 if self.adam_mode == "dense" else 1
        p.dropout_p = m.user_tower.dropout_p if m.training else 0.0
        p.seed = self.seed
        p.keep_mask_user = p.keep_mask_pos = p.keep_mask_neg = None
        p.padding_idx = 0
        p.loss, p.err_flag = ptr(self.loss_dev), ptr(self.err_flag)
        p.workspace, p.workspace_bytes = ptr(self.ws), self.ws.numel()
        p.next_batch = C.addressof(self._sampler[0]) if self._sampler is not None else None
        return p

    def _launch(self, masks=None) -> None:
        if not self._is_flat():
            self._flatten()
            self._graph = None
        p = self._make_params()
        if masks is not None:
            self._mask_keep = [None if t is None else t.to(self.dev, torch.uint8).contiguous() for t in masks]
            p.keep_mask_user, p.keep_mask_pos, p.keep_mask_neg = (ptr(t) for t in self._mask_keep)
        self._params = p
        check(self.lib.rb200_bpr_step(C.byref(p), stream_ptr()), "rb200_bpr_step")

    def step(self, masks=None) -> torch.Tensor:
        """One optimiser step on the batch currently in the staging buffers (``user_ids`` … ``neg_extra``).
        Returns the device scalar holding the loss (no synchronisation)."""
        if self._B is None:
            raise RB200Error("no batch staged: call load_batch()/step_host() first")
        with torch.cuda.device(self.dev):
            training_p = self.model.user_tower.dropout_p if self.model.training else 0.0
            key = (self._B, training_p)
            if not self.use_graph or masks is not None:
                self._launch(masks)
            elif self._warm_key != key:
                self._launch()                          # first step of a shape runs eagerly: sets kernel attributes
                self._warm_key, self._graph = key, None
            elif self._graph is None or not self._is_flat():
                if not self._is_flat():
                    self._flatten()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):               # capture enqueues nothing; replay() below runs the step
                    self._launch()
                self._graph = g
                g.replay()
            else:
                self._graph.replay()
        self._steps_done += 1
        return self.loss_dev

    _warm_key = None

    def load_batch(self, user_ids, pos_ids, pos_genres=None, neg_ids=None, neg_genres=None) -> None:
        """Stage a batch that already lives on the device (device→device copies into the static buffers)."""
        B = user_ids.numel()
        if self._B != B:
            self._alloc(B)
        self.user_ids.copy_(user_ids)
        self.pos_ids.copy_(pos_ids)
        if neg_ids is not None:
            self.neg_ids.copy_(neg_ids)
        if self.item_extra_table is None and self.E:
            self.pos_extra.copy_(pos_genres)
            if neg_genres is not None:
                self.neg_extra.copy_(neg_genres)

    def load_packed(self, packed: torch.Tensor) -> None:
        """Stage a packed batch (``pack_host`` layout) from pinned host or device memory with one async copy."""
        self.batch_dev.copy_(packed, non_blocking=True)

    def step_host(self, user_ids, pos_ids, pos_genres, neg_ids, neg_genres) -> float:
        """End-to-end step on a host batch: pack → one H2D copy → graph replay → D2H of the loss (synchronises,
        like ``loss.item()`` at train_embeddings.py:194)."""
        self.load_packed(self.pack_host(user_ids, pos_ids, pos_genres, neg_ids, neg_genres))
        return float(self.step().item())

    def train_epoch(self, packed_batches) -> float:
        """The reference's ``train_epoch`` loop (train_embeddings.py:170-199) over an iterable of PINNED packed host batches
        (``pack_host(...).pin_memory()``), pipelined: the host→device copy of batch i+1 runs on a copy stream while batch i
        computes, and each step's loss is copied back asynchronously into a pinned array (one 4-byte D2H per step) that
        is summed once at the end — the same per-batch traffic as ``step_host`` without a host synchronisation per step.
        Returns the mean loss over the batches, like the reference."""
        with torch.cuda.device(self.dev):
            main = torch.cuda.current_stream(self.dev)
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(self.dev)
                self._stage = [None, None]
                self._stage_ready = [torch.cuda.Event(), torch.cuda.Event()]
                self._stage_free = [torch.cuda.Event(), torch.cuda.Event()]
            it = iter(packed_batches)
            losses_host = None
            n = 0

            def prefetch(k, batch):
                if self._stage[k] is None or self._stage[k].shape != batch.shape:
                    self._stage[k] = torch.empty(batch.shape, dtype=batch.dtype, device=self.dev)
                with torch.cuda.stream(self._copy_stream):
                    self._copy_stream.wait_event(self._stage_free[k])
                    self._stage[k].copy_(batch, non_blocking=True)
                    self._stage_ready[k].record(self._copy_stream)

            for k in (0, 1):
                self._stage_free[k].record(main)
            nxt = next(it, None)
            if nxt is not None:
                prefetch(0, nxt)
            while nxt is not None:
                k = n & 1
                cur, nxt = nxt, next(it, None)
                if nxt is not None:
                    prefetch(k ^ 1, nxt)
                main.wait_event(self._stage_ready[k])
                self.batch_dev.copy_(self._stage[k], non_blocking=True)       # device→device into the graph's static buffer
                self._stage_free[k].record(main)
                loss = self.step()
                if losses_host is None or n >= losses_host.numel():
                    grown = torch.zeros(max(1024, 2 * n), dtype=torch.float32).pin_memory()
                    if losses_host is not None:
                        main.synchronize()
                        grown[:n] = losses_host[:n]
                    losses_host = grown
                losses_host[n:n + 1].copy_(loss, non_blocking=True)
                n += 1
            main.synchronize()
        return float(losses_host[:n].double().mean()) if n else float("nan")

    def check_ids(self) -> None:
        if int(self.err_flag.item()) != 0:
            raise IndexError("an id in a previous batch was outside its embedding table (torch would raise "
                             "'index out of range in self')")

    def views(self) -> Dict[str, torch.Tensor]:
        """Gradients of the last step (for tests): flat MLP grads and compact unique-row gradients."""
        v = StepViews()
        check(self.lib.rb200_bpr_step_views(C.byref(self._params), C.byref(v)), "rb200_bpr_step_views")
        torch.cuda.synchronize(self.dev)

        def wrap(addr, n, dtype):
            out = torch.empty(n, dtype=dtype, device=self.dev)
            size = n * out.element_size()
            off = addr - self.ws.data_ptr()
            out.view(torch.uint8).copy_(self.ws[off:off + size])
            return out

        B, D = self._B, self.D
        items = 2 if self.loss_kind == 0 else 1
        nu = int(wrap(v.user_n_uniq, 1, torch.int32).item())
        ni = int(wrap(v.item_n_uniq, 1, torch.int32).item())
        return {
            "user_mlp_grad": wrap(v.user_mlp_grad, self._user_flat.numel(), torch.float32),
            "item_mlp_grad": wrap(v.item_mlp_grad, self._item_flat.numel(), torch.float32),
            "user_uniq_ids": wrap(v.user_uniq_ids, B, torch.int64)[:nu],
            "item_uniq_ids": wrap(v.item_uniq_ids, items * B, torch.int64)[:ni],
            "user_uniq_grads": wrap(v.user_uniq_grads, B * D, torch.float32).view(B, D)[:nu],
            "item_uniq_grads": wrap(v.item_uniq_grads, items * B * D, torch.float32).view(items * B, D)[:ni],
            "user_emb": wrap(v.user_emb, B * D, torch.float32).view(B, D),
            "pos_emb": wrap(v.pos_emb, B * D, torch.float32).view(B, D),
        }

    # ---- checkpointing of the optimiser (the reference saves none; offered for resume) ------------ #
    def state_dict(self) -> Dict:
        st = self._pull_opt()
        return {"moments": {k: v.detach().cpu() for k, v in self.state.items()}, "step": int(st.step), "lr": float(st.lr)}

    def load_state_dict(self, sd: Dict) -> None:
        for k, v in sd["moments"].items():
            self.state[k].copy_(v)
        st = self._pull_opt()
        st.step, st.lr = int(sd["step"]), float(sd["lr"])
        self._opt_host = st
        self._push_opt()


class DataParallelBPRTrainer(FusedBPRTrainer):
    """Data-parallel form of the fused step for tables that fit every GPU (the ML-1M shape, BASELINE C1/C2): every rank
    holds a replica, processes its own batch, and the replicas are kept identical by ONE all-reduce of the dense gradients
    (MLPs + both tables, 2.7 MB at the reference's sizes) between the two halves of the step:

        rb200_bpr_step (… → dense gradients in `dp_grads`, loss gradient scaled by 1/world)
        torch.distributed.all_reduce(dp_grads)                                   # NCCL over NVLink (allreduce="nccl"), or
        barrier; rb200_allreduce_twoshot; barrier                                # peer memory, no NCCL (allreduce="p2p")
        rb200_bpr_apply (Σg² → clip → Adam on every parameter, as torch.optim.Adam on dense grads does)

    — all three captured in ONE CUDA graph after two eager steps (call ``close()`` before ``destroy_process_group()``).

    The result equals the single-process step on the concatenated global batch (mean loss over world·B samples).
    Parameters are broadcast from rank 0 at construction.  Huge tables use ``sharded.ShardedBPRTrainer`` instead."""

    def __init__(self, model: TwoTowerModel, group=None, allreduce: str = "nccl", **kw):
        import torch.distributed as dist
        if allreduce not in ("nccl", "p2p"):
            raise ValueError("allreduce must be 'nccl' or 'p2p'")
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        kw.setdefault("adam_mode", "dense")
        if kw["adam_mode"] != "dense":
            raise ValueError("data-parallel replicas use dense Adam (the all-reduced gradients are dense)")
        seed = kw.pop("seed", None)
        base_seed = torch.initial_seed() if seed is None else seed
        super().__init__(model, seed=base_seed + 7919 * self.rank, **kw)     # independent dropout masks per rank
        if self.world > 1:
            for p in model.parameters():
                dist.broadcast(p.data, src=0, group=group)
            self._flatten()
        n = self.lib.rb200_bpr_dp_grad_floats(self.D, self.H, self.E, model.n_users + 1, model.n_items + 1)
        # dense gradients + one trailing float for the loss (pre-scaled by 1/world on the device): ONE all-reduce per step
        self._dp_hdl, self._dp_ptrs, self._dp_mc = None, None, 0
        n_buf = (n + 4 + 3) // 4 * 4
        if self.world > 1 and allreduce == "p2p":
            # the buffer lives in symmetric memory (one NVLink / NVSwitch box): the all-reduce is a two-shot kernel over peer memory
            # between two cross-GPU barriers (rb200_allreduce_twoshot) instead of ncclAllReduce
            import torch.distributed._symmetric_memory as symm_mem
            self._dp_buf = symm_mem.empty(n_buf, dtype=torch.float32, device=self.dev)
            self._dp_buf.zero_()
            self._dp_hdl = symm_mem.rendezvous(self._dp_buf, group if group is not None else dist.group.WORLD)
            self._dp_ptrs = [int(p) for p in self._dp_hdl.buffer_ptrs]
            # NVLS: the reduction inside the switch when the fabric offers a multicast object (RB200_DP_MULTIMEM=0: two-shot kernel)
            import os
            mc = int(getattr(self._dp_hdl, "multicast_ptr", 0) or 0)
            self._dp_mc = mc if os.environ.get("RB200_DP_MULTIMEM", "1") != "0" else 0
        else:
            self._dp_buf = torch.zeros(n_buf, dtype=torch.float32, device=self.dev)
        self.dp_grads = self._dp_buf[:n]
        self.loss_sum = self._dp_buf[n:n + 1]
        self._eager_dp = 0

    def close(self) -> None:
        """Release the captured step: a CUDA graph that holds NCCL nodes must be gone before ``destroy_process_group()``."""
        self._graph = None

    def _make_params(self) -> StepParams:
        p = super()._make_params()
        p.grad_scale = 1.0 / self.world
        p.dp_grads = ptr(self.dp_grads)
        p.loss = ptr(self.loss_sum)
        return p

    def _phase(self, which: int) -> None:
        if not self._is_flat():
            self._flatten()
            self._graph = None
        p = self._make_params()
        self._params = p
        fn = self.lib.rb200_bpr_step if which == 0 else self.lib.rb200_bpr_apply
        check(fn(C.byref(p), stream_ptr()), "rb200_bpr_step" if which == 0 else "rb200_bpr_apply")

    def _whole(self) -> None:
        self._phase(0)
        if self.world > 1:
            if self._dp_hdl is not None:
                self._dp_hdl.barrier(channel=0, timeout_ms=60000)          # every rank's gradients are complete
                if self._dp_mc:
                    check(self.lib.rb200_allreduce_multimem(self._dp_mc, self.world, self.rank, self._dp_buf.numel(), stream_ptr()),
                          "rb200_allreduce_multimem")
                else:
                    check(self.lib.rb200_allreduce_twoshot((C.c_void_p * self.world)(*self._dp_ptrs), self.world, self.rank,
                                                           self._dp_buf.numel(), stream_ptr()), "rb200_allreduce_twoshot")
                self._dp_hdl.barrier(channel=1, timeout_ms=60000)          # every slice has been delivered everywhere
            else:
                self.dist.all_reduce(self._dp_buf, group=self.group)
        self._phase(1)

    def step(self, masks=None) -> torch.Tensor:
        """Returns the device scalar holding the GLOBAL mean loss."""
        if self._B is None:
            raise RB200Error("no batch staged: call load_batch()/step_host() first")
        with torch.cuda.device(self.dev):
            key = (self._B, self.model.user_tower.dropout_p if self.model.training else 0.0)
            use_graph = self.use_graph and self._warm_key == key and self._eager_dp >= 2
            if use_graph and self._graph is None:
                # the whole step — both halves AND the NCCL all-reduce between them — is ONE graph (round 1 replayed two graphs
                # with an eagerly issued all-reduce in between: three host launches and a serial collective per step)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._whole()
                self._graph = g
            if use_graph:
                self._graph.replay()
            else:
                self._whole()             # two eager steps first: communicator, kernel attributes and allocator warm up
                self._eager_dp = self._eager_dp + 1 if self._warm_key == key else 1
            self._warm_key = key
        self._steps_done += 1
        return self.loss_sum
