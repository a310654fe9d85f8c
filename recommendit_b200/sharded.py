"""Row-sharded two-tower training and sharded retrieval over one NVLink/NVSwitch box (BASELINE configs C4 and C5).

The reference is single-process; this is the scaled form of the same step: one process per GPU,
``torch.distributed`` (NCCL) for the plumbing, the kernels of ``librb200.so`` for the arithmetic.

Training step on rank r, local batch of B samples, tables sharded by ``owner(id) = id % world``,
``local_row(id) = id // world`` (modulo sharding flattens Zipf-skewed id popularity):

  1. bucket the batch's ids by owner (stable)                         → all-to-all: ids to their owners
  2. owners gather the requested rows (``rb200_gather_rows``)          → all-to-all: rows back
  3. towers fwd → ``bpr_loss`` → towers bwd on the local batch; the received row buffer acts as the "embedding
     table" and the inverse bucket permutation as the "ids", so no un-permute copy is needed
  4. row gradients in bucket order                                    → all-to-all: gradients to the owners
  5. owners: deterministic sorted-segment sum per local row; MLP gradients all-reduced (sum; every rank scales its
     loss gradient by 1/world so the sum is the gradient of the global-batch mean)
  6. Σg² partials + loss all-reduced → global clip coefficient → Adam on the local shards and on the replicated MLPs

The global result is independent of ``world`` up to fp32 summation order (tests/test_sharded_gloo.py checks it against
the single-process oracle on the concatenated batch).

Retrieval (C5): database rows sharded contiguously, queries replicated, per-shard ``rb200_flat_search`` top-k,
all-gather of ``(score, id)[nq, k]``, ``rb200_topk_merge``.

``ops`` is the arithmetic back end.  The product uses ``CudaOps`` (C ABI).  It is a parameter only so that the
host-side exchange logic can be exercised on CPU/gloo in the tests with a stand-in built on the oracle.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib
from ._lib import OptState, SumsqSeg, TowerBwdJob, TowerJob, check, ptr, stream_ptr, workspace


# --------------------------------------------------------------------------------------------------------- #
# exchange plan (pure index arithmetic: runs wherever the tensors live)
# --------------------------------------------------------------------------------------------------------- #
@dataclass
class Route:
    perm: torch.Tensor          # [n] bucket order → sample index (stable by owner)
    inv: torch.Tensor           # [n] sample index → position in bucket order
    local_rows: torch.Tensor    # [n] local row of each request, bucket order
    send_counts: List[int]      # requests per owner
    recv_counts: List[int] = None
    recv_rows: torch.Tensor = None   # local rows requested from this rank, grouped by source


def make_route(ids: torch.Tensor, world: int) -> Route:
    owner = ids % world
    perm = torch.argsort(owner, stable=True)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(ids.numel(), device=ids.device, dtype=perm.dtype)
    counts = torch.bincount(owner, minlength=world)
    return Route(perm=perm, inv=inv, local_rows=torch.div(ids, world, rounding_mode="floor")[perm].contiguous(),
                 send_counts=[int(c) for c in counts.tolist()])


def route_reference(user_ids: torch.Tensor, item_ids: torch.Tensor, world: int, nu_by_rank: torch.Tensor):
    """What ``rb200_route_plan`` computes, restated with generic tensor ops — used by the CPU test back end
    (tests/) and as the checker of the kernel in its GPU test; the product path calls the kernel."""
    ids = torch.cat([user_ids, item_ids])
    owner = ids % world
    local = torch.div(ids, world, rounding_mode="floor")
    local[user_ids.numel():] += nu_by_rank[owner[user_ids.numel():]]          # item rows sit behind the owner's user rows
    perm = torch.argsort(owner, stable=True)
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(ids.numel(), device=ids.device, dtype=perm.dtype)
    return perm, inv, local[perm].contiguous(), torch.bincount(owner, minlength=world)


def route_padded_reference(user_ids: torch.Tensor, item_ids: torch.Tensor, world: int, nu_by_rank: torch.Tensor, capacity: int,
                           overflow: torch.Tensor):
    """What ``rb200_route_plan_padded`` computes, restated with generic tensor ops (CPU test back end and the checker of the
    kernel in its GPU test): → (slot of each sample-order request [n], rows per slot [world·capacity], -1 = empty)."""
    W, C = world, capacity
    perm, inv, local_rows, send = route_reference(user_ids, item_ids, W, nu_by_rank)
    n = perm.numel()
    ends = torch.cumsum(send, 0)
    pos = torch.arange(n, device=perm.device, dtype=torch.int64)
    owner = torch.bucketize(pos, ends, right=True).clamp_(max=W - 1)           # bucket of each position of the bucket order
    off = pos - (ends - send)[owner]
    fits = off < C
    overflow += (~fits).sum()
    slot = torch.where(fits, owner * C + off, torch.full_like(pos, W * C))     # what does not fit goes to the dummy slot W·C
    send_rows = torch.full((W * C + 1,), -1, dtype=torch.int64, device=perm.device)
    send_rows.index_copy_(0, slot, local_rows)
    return slot[inv], send_rows[:W * C].contiguous()


def shard_rows(n_rows: int, world: int, rank: int) -> int:
    """rows of a table with global ids 0..n_rows-1 owned by ``rank`` under modulo sharding"""
    return (n_rows - rank + world - 1) // world if n_rows > rank else 0


def all_to_all_counts(send_counts: List[int], group, device) -> List[int]:
    s = torch.tensor(send_counts, dtype=torch.int64, device=device)
    r = torch.empty_like(s)
    dist.all_to_all_single(r, s, group=group)
    return [int(c) for c in r.tolist()]


def all_to_all_var(send: torch.Tensor, send_counts: List[int], recv_counts: List[int], group) -> torch.Tensor:
    out = torch.empty((sum(recv_counts),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
    dist.all_to_all_single(out, send.contiguous(), output_split_sizes=recv_counts, input_split_sizes=send_counts, group=group)
    return out


# positions of rb200_opt_state fields (include/rb200.h) in float64 / float32 views of the 128-byte state
_OPT_SUMSQ_F64 = OptState.sumsq.offset // 8
_OPT_LOSS_F32 = OptState.loss.offset // 4


# --------------------------------------------------------------------------------------------------------- #
# arithmetic back end over the C ABI
# --------------------------------------------------------------------------------------------------------- #
class CudaOps:
    """Thin, allocation-only wrappers over ``librb200.so`` (device tensors in, device tensors out)."""

    def __init__(self):
        self.lib = _lib.load()

    def mode(self, D, H, jobs) -> int:
        from .two_tower import tower_mode_for
        E = max(0 if j.get("extra") is None else j["extra"].shape[1] for j in jobs)
        return tower_mode_for(D, H, E)

    def route(self, user_ids: torch.Tensor, item_ids: torch.Tensor, world: int, nu_by_rank: torch.Tensor):
        """→ (perm, inv, local_rows, send_counts) of the combined request list, one C call (``rb200_route_plan``)"""
        n = user_ids.numel() + item_ids.numel()
        dev = user_ids.device
        perm = torch.empty(n, dtype=torch.int64, device=dev)
        inv, local = torch.empty_like(perm), torch.empty_like(perm)
        send = torch.empty(world, dtype=torch.int64, device=dev)
        wsb = self.lib.rb200_route_plan_workspace_bytes(n, world)
        ws = workspace(wsb, dev)
        check(self.lib.rb200_route_plan(ptr(user_ids), user_ids.numel(), ptr(item_ids), item_ids.numel(), world, ptr(nu_by_rank),
                                        ptr(perm), ptr(inv), ptr(local), ptr(send), ptr(ws), wsb, stream_ptr()), "rb200_route_plan")
        return perm, inv, local, send

    def route_padded(self, user_ids: torch.Tensor, item_ids: torch.Tensor, world: int, nu_by_rank: torch.Tensor, capacity: int,
                     overflow: torch.Tensor):
        """→ (slot of each sample-order request [n], rows per slot [world·capacity], -1 = empty), one C call
        (``rb200_route_plan_padded``); requests beyond a bucket's capacity are added to ``overflow``"""
        n = user_ids.numel() + item_ids.numel()
        dev = user_ids.device
        slot = torch.empty(n, dtype=torch.int64, device=dev)
        send_rows = torch.empty(world * capacity, dtype=torch.int64, device=dev)
        wsb = self.lib.rb200_route_plan_padded_workspace_bytes(n, world)
        ws = workspace(wsb, dev)
        check(self.lib.rb200_route_plan_padded(ptr(user_ids), user_ids.numel(), ptr(item_ids), item_ids.numel(), world, ptr(nu_by_rank),
                                               capacity, ptr(slot), ptr(send_rows), ptr(overflow), ptr(ws), wsb, stream_ptr()),
              "rb200_route_plan_padded")
        return slot, send_rows

    def gather_rows(self, table: torch.Tensor, rows: torch.Tensor) -> torch.Tensor:
        out = torch.empty(rows.numel(), table.shape[1], dtype=torch.float32, device=table.device)
        if rows.numel():
            check(self.lib.rb200_gather_rows(ptr(table), ptr(rows), rows.numel(), table.shape[1], table.shape[0], ptr(out),
                                             stream_ptr()), "rb200_gather_rows")
        return out

    def gather_rows_sharded(self, shard_ptrs: List[int], user_rows_by_rank: List[int], user_ids: torch.Tensor, item_ids: torch.Tensor,
                            n_user_rows: int, n_item_rows: int, D: int, err_flag: Optional[torch.Tensor],
                            out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """out[r] = the row of request r (sample order [users | items]) read straight from its owner's shard through the peer
        pointers (``rb200_gather_rows_sharded``): no id exchange, no owner-side gather, no row exchange"""
        W = len(shard_ptrs)
        n = user_ids.numel() + item_ids.numel()
        if out is None:
            out = torch.empty(n, D, dtype=torch.float32, device=user_ids.device)
        check(self.lib.rb200_gather_rows_sharded((C.c_void_p * W)(*shard_ptrs), (C.c_int64 * W)(*user_rows_by_rank), W, ptr(user_ids),
                                                 user_ids.numel(), ptr(item_ids), item_ids.numel(), n_user_rows, n_item_rows, D, ptr(out),
                                                 ptr(err_flag), stream_ptr()), "rb200_gather_rows_sharded")
        return out

    def push_rows_sharded(self, grad_ptrs: List[int], row_ptrs: List[int], rank: int, capacity: int, drows: torch.Tensor,
                          slot: torch.Tensor, send_rows: Optional[torch.Tensor]) -> None:
        """gradient rows (+ the plan's row list unless ``send_rows`` is None: it then went ahead with :meth:`push_row_lists`) → the
        owners' receive buckets through the peer pointers (``rb200_push_rows_sharded``)"""
        W = len(grad_ptrs)
        check(self.lib.rb200_push_rows_sharded((C.c_void_p * W)(*grad_ptrs), (C.c_void_p * W)(*row_ptrs), W, rank, capacity, ptr(drows),
                                               ptr(slot), drows.shape[0], drows.shape[1], ptr(send_rows), stream_ptr()),
              "rb200_push_rows_sharded")

    def push_row_lists(self, row_ptrs: List[int], rank: int, capacity: int, send_rows: torch.Tensor) -> None:
        """the plan's row list → the owners' row buckets (``rb200_push_row_lists_sharded``)"""
        W = len(row_ptrs)
        check(self.lib.rb200_push_row_lists_sharded((C.c_void_p * W)(*row_ptrs), W, rank, capacity, ptr(send_rows), stream_ptr()),
              "rb200_push_row_lists_sharded")

    def allreduce_oneshot(self, src_ptrs: List[int], n: int, out: torch.Tensor) -> None:
        W = len(src_ptrs)
        check(self.lib.rb200_allreduce_oneshot((C.c_void_p * W)(*src_ptrs), W, n, ptr(out), stream_ptr()), "rb200_allreduce_oneshot")

    def scalars_publish(self, opt: torch.Tensor, loss: torch.Tensor, scale: float, slot: torch.Tensor) -> None:
        check(self.lib.rb200_sharded_scalars_publish(ptr(opt), ptr(loss), scale, ptr(slot), stream_ptr()), "rb200_sharded_scalars_publish")

    def scalars_finish(self, slot_ptrs: List[int], dense_grad: torch.Tensor, opt: torch.Tensor) -> None:
        """scalars_reduce + Σ dense_grad² + the clip coefficient in one launch (``rb200_sharded_scalars_finish``)"""
        W = len(slot_ptrs)
        check(self.lib.rb200_sharded_scalars_finish((C.c_void_p * W)(*slot_ptrs), W, ptr(dense_grad), dense_grad.numel(), ptr(opt),
                                                    stream_ptr()), "rb200_sharded_scalars_finish")

    def adam_rows_dense2(self, w, m, v, uniq, ug, nu, dense, opt) -> None:
        """Adam on the touched rows of ``w`` and on the dense blocks ``dense = [(w, g, m, v), …]`` (at most two) in one launch"""
        d = list(dense) + [(None, None, None, None)] * (2 - len(dense))
        n = [0 if t[0] is None else t[0].numel() for t in d]
        check(self.lib.rb200_adam_rows_dense2(ptr(w), ptr(m), ptr(v), w.shape[1], ptr(uniq), ptr(ug), ptr(nu), uniq.numel(),
                                              ptr(d[0][0]), ptr(d[0][1]), ptr(d[0][2]), ptr(d[0][3]), n[0],
                                              ptr(d[1][0]), ptr(d[1][1]), ptr(d[1][2]), ptr(d[1][3]), n[1], ptr(opt), stream_ptr()),
              "rb200_adam_rows_dense2")

    def scalars_reduce(self, slot_ptrs: List[int], opt: torch.Tensor) -> None:
        W = len(slot_ptrs)
        check(self.lib.rb200_sharded_scalars_reduce((C.c_void_p * W)(*slot_ptrs), W, ptr(opt), stream_ptr()), "rb200_sharded_scalars_reduce")

    def towers_fwd(self, jobs: List[dict], D: int, H: int, drop_p: float, seed: int, offset: int, offset_dev=None) -> None:
        """``offset_dev``: device int64 added (× 3) to the dropout offset inside the kernel — the optimizer's step counter, so
        that replays of a captured step draw fresh masks"""
        arr = (TowerJob * len(jobs))()
        for i, j in enumerate(jobs):
            arr[i] = TowerJob(ptr(j["table"]), ptr(j["ids"]), ptr(j.get("extra")), ptr(j["W1"]), ptr(j["b1"]), ptr(j["W2"]),
                              ptr(j["b2"]), ptr(j["out"]), ptr(j["hid"]), ptr(j["denom"]), None, j["table"].shape[0],
                              j["ids"].numel(), 0 if j.get("extra") is None else j["extra"].shape[1], 0, None)
        mode = self.mode(D, H, jobs)
        E = max(0 if j.get("extra") is None else j["extra"].shape[1] for j in jobs)
        wsb = self.lib.rb200_tower_fwd_workspace_bytes(len(jobs), D, H, E, mode)
        ws = workspace(wsb, jobs[0]["out"].device) if wsb else None
        check(self.lib.rb200_tower_fwd(arr, len(jobs), D, H, drop_p, seed, offset, offset_dev, mode, None, ptr(ws), wsb, stream_ptr()),
              "rb200_tower_fwd")

    def bpr_pair(self, u, p, n, grad_scale: float):
        B, D = u.shape
        loss = torch.empty(1, dtype=torch.float32, device=u.device)
        du, dp, dn = torch.empty_like(u), torch.empty_like(p), torch.empty_like(n)
        wsb = self.lib.rb200_bpr_pair_workspace_bytes(B)
        ws = workspace(wsb, u.device)
        check(self.lib.rb200_bpr_pair(ptr(u), ptr(p), ptr(n), B, D, ptr(loss), ptr(du), ptr(dp), ptr(dn), grad_scale, ptr(ws), wsb,
                                      stream_ptr()), "rb200_bpr_pair")
        return loss, du, dp, dn

    def towers_bwd(self, jobs: List[dict], D: int, H: int, drop_p: float, grads_out: torch.Tensor) -> None:
        arr = (TowerBwdJob * len(jobs))()
        E = 0 if jobs[0].get("extra") is None else jobs[0]["extra"].shape[1]
        for i, j in enumerate(jobs):
            arr[i] = TowerBwdJob(ptr(j["table"]), ptr(j["ids"]), ptr(j.get("extra")), j["table"].shape[0], j["ids"].numel(), E, 0,
                                 ptr(j["W1"]), ptr(j["W2"]), ptr(j["dY"]), ptr(j["out"]), ptr(j["denom"]), ptr(j["hid"]),
                                 ptr(j["dpre"]), ptr(j["dact"]), ptr(j["dRows"]), None)
        wsb = self.lib.rb200_tower_bwd_workspace_bytes(D, H, E)
        ws = workspace(wsb, grads_out.device)
        check(self.lib.rb200_tower_bwd(arr, len(jobs), D, H, drop_p, self.mode(D, H, jobs), ptr(grads_out), 0, ptr(ws), wsb, stream_ptr()),
              "rb200_tower_bwd")

    def scatter_rows(self, ids: torch.Tensor, rows: torch.Tensor, n_rows: int, padding_row: int):
        """→ (uniq_ids [cap], uniq_grads [cap, D], n_uniq [1] int32), deterministic; ``padding_row`` < 0 ⇒ none"""
        B, D = rows.shape
        cap = max(B, 1)
        uniq = torch.empty(cap, dtype=torch.int64, device=rows.device)
        ug = torch.empty(cap, D, dtype=torch.float32, device=rows.device)
        nu = torch.zeros(1, dtype=torch.int32, device=rows.device)
        if B:
            wsb = self.lib.rb200_scatter_workspace_bytes(B, n_rows)
            ws = workspace(wsb, rows.device)
            check(self.lib.rb200_scatter_rows(ptr(ids), ptr(rows), B, D, n_rows, padding_row, None, ptr(uniq), ptr(ug), ptr(nu), None,
                                              ptr(ws), wsb, stream_ptr()), "rb200_scatter_rows")
        return uniq, ug, nu

    def scatter_plan(self, ids: torch.Tensor, n_rows: int, padding_row: int):
        """first phase of :meth:`scatter_rows` (needs only the ids) → (uniq_ids [cap], n_uniq [1] int32, workspace)"""
        B = ids.numel()
        uniq = torch.empty(max(B, 1), dtype=torch.int64, device=ids.device)
        nu = torch.zeros(1, dtype=torch.int32, device=ids.device)
        wsb = self.lib.rb200_scatter_workspace_bytes(B, n_rows)
        ws = torch.empty(wsb, dtype=torch.uint8, device=ids.device)        # (not the shared scratch: it has to survive until apply)
        check(self.lib.rb200_scatter_plan(ptr(ids), B, n_rows, padding_row, ptr(uniq), ptr(nu), None, ptr(ws), wsb, stream_ptr()),
              "rb200_scatter_plan")
        return uniq, nu, ws

    def scatter_apply(self, rows: torch.Tensor, n_rows: int, uniq: torch.Tensor, nu: torch.Tensor, ws: torch.Tensor) -> torch.Tensor:
        """second phase → uniq_grads [cap, D]"""
        B, D = rows.shape
        ug = torch.empty(max(B, 1), D, dtype=torch.float32, device=rows.device)
        check(self.lib.rb200_scatter_apply(ptr(rows), B, D, n_rows, None, ptr(uniq), ptr(ug), ptr(nu), ptr(ws), ws.numel(), stream_ptr()),
              "rb200_scatter_apply")
        return ug

    def sumsq(self, opt: torch.Tensor, segs: List[Tuple[torch.Tensor, Optional[torch.Tensor], int]]) -> None:
        """opt.sumsq += Σ x² ; a segment is (tensor, count tensor or None, row_len)"""
        wsb = self.lib.rb200_sumsq_workspace_bytes()
        ws = workspace(wsb, opt.device)
        for i in range(0, len(segs), 4):
            chunk = segs[i:i + 4]
            arr = (SumsqSeg * 4)()
            for k, (t, cnt, rl) in enumerate(chunk):
                arr[k] = SumsqSeg(ptr(t), t.numel(), ptr(cnt), rl)
            check(self.lib.rb200_sumsq_accumulate(ptr(opt), arr, len(chunk), ptr(ws), wsb, stream_ptr()), "rb200_sumsq_accumulate")

    def begin_step(self, opt: torch.Tensor) -> None:
        check(self.lib.rb200_opt_begin_step(ptr(opt), stream_ptr()), "rb200_opt_begin_step")

    def grad_norm_clip(self, opt: torch.Tensor) -> None:
        """total_norm = sqrt(opt.sumsq); clip_coef = min(1, max_norm / (total_norm + 1e-6)) — on the device"""
        check(self.lib.rb200_grad_norm_clip(ptr(opt), stream_ptr()), "rb200_grad_norm_clip")

    def adam_dense(self, w, g, m, v, opt) -> None:
        check(self.lib.rb200_adam_dense(ptr(w), ptr(g), ptr(m), ptr(v), w.numel(), ptr(opt), stream_ptr()), "rb200_adam_dense")

    def adam_rows(self, w, m, v, uniq, ug, nu, opt) -> None:
        check(self.lib.rb200_adam_rows(ptr(w), ptr(m), ptr(v), w.shape[1], ptr(uniq), ptr(ug), ptr(nu), uniq.numel(), ptr(opt),
                                       stream_ptr()), "rb200_adam_rows")

    def adam_table_dense(self, w, m, v, uniq, ug, nu, slot, opt) -> None:
        """dense (reference-exact) mode: slots are set from the compact list, every row updated, slots reset"""
        check(self.lib.rb200_scatter_set_slots(ptr(uniq), ptr(nu), uniq.numel(), ptr(slot), stream_ptr()), "rb200_scatter_set_slots")
        check(self.lib.rb200_adam_table_dense(ptr(w), ptr(m), ptr(v), w.shape[0], w.shape[1], ptr(slot), ptr(ug), ptr(opt),
                                              stream_ptr()), "rb200_adam_table_dense")
        check(self.lib.rb200_scatter_reset_slots(ptr(uniq), ptr(nu), uniq.numel(), ptr(slot), stream_ptr()), "rb200_scatter_reset_slots")

    # host <-> device mirror of rb200_opt_state
    def read_opt(self, opt: torch.Tensor) -> OptState:
        return OptState.from_buffer_copy(bytes(opt.cpu().numpy().tobytes()))

    def write_opt(self, opt: torch.Tensor, st: OptState) -> None:
        opt.copy_(torch.frombuffer(bytearray(bytes(st)), dtype=torch.uint8))


# --------------------------------------------------------------------------------------------------------- #
# the sharded trainer
# --------------------------------------------------------------------------------------------------------- #
class ShardedBPRTrainer:
    """Two-tower BPR training with row-sharded embedding tables (BASELINE config C4)."""

    def __init__(self, n_users: int, n_items: int, embed_dim: int = 128, hidden_dim: int = 128, n_genres: int = 18,
                 lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-5, max_norm: float = 1.0,
                 adam_mode: str = "rows", device=None, group=None, ops=None, seed: int = 0, init: Optional[Dict] = None,
                 exchange: str = "exact", capacity_factor: float = 2.0, use_cuda_graph: bool = False, dropout: float = 0.0,
                 check_every: int = 256):
        """``dropout``: the towers' dropout probability (the reference's model default is 0.1; 0.0 here keeps the parity runs
        mask-free).  Masks are drawn in-kernel (Philox) from (seed + 7919·rank, optimizer step).  ``check_every``: padded
        exchange — every that many steps the overflow counter is read back and an overflow raises (0 = only on
        ``check_exchange()``)."""
        if exchange not in ("exact", "padded", "p2p"):
            raise ValueError("exchange must be 'exact' (variable-size all-to-alls, one host sync per step), 'padded' "
                             "(fixed-capacity all-to-alls: no host sync, CUDA-graph capturable) or 'p2p' (rows read from / gradients "
                             "written to the owners' memory over NVLink: no collective on the row path, CUDA-graph capturable)")
        if use_cuda_graph and exchange == "exact":
            raise ValueError("use_cuda_graph needs exchange='padded' or 'p2p' (the exact exchange reads its split sizes on the host)")
        if not 0.0 <= dropout < 1.0:
            raise ValueError("dropout must be in [0, 1)")
        self.exchange, self.capacity_factor, self.use_graph = exchange, float(capacity_factor), use_cuda_graph
        self.dropout, self.check_every, self.seed = float(dropout), int(check_every), int(seed)
        self._graph, self._static_in, self._static_loss, self._eager_steps = None, None, None, 0
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.ops = ops if ops is not None else CudaOps()
        self.D, self.H, self.E = embed_dim, hidden_dim, n_genres
        self.n_user_rows, self.n_item_rows = n_users + 1, n_items + 1
        self.adam_mode = adam_mode
        f32 = dict(dtype=torch.float32, device=self.dev)
        W, r = self.world, self.rank
        nu, ni = shard_rows(self.n_user_rows, W, r), shard_rows(self.n_item_rows, W, r)
        # Both shards live in ONE tensor (user rows first, then item rows): a request is a row of the combined shard, so
        # one exchange / gather / scatter / Adam launch serves both tables.  user_table / item_table are views.
        self._symm, self._tab_hdl, self._buckets = None, None, {}
        self._nu_host = [shard_rows(self.n_user_rows, W, k) for k in range(W)]
        if exchange == "p2p" and W > 1:
            # the shard lives in symmetric memory: every rank maps every other rank's shard (NVLink peer access through NVSwitch)
            import torch.distributed._symmetric_memory as symm_mem
            self._symm = symm_mem
            rows_alloc = shard_rows(self.n_user_rows, W, 0) + shard_rows(self.n_item_rows, W, 0)        # the largest shard
            flat = symm_mem.empty(rows_alloc * embed_dim, dtype=torch.float32, device=self.dev)
            self._tab_hdl = symm_mem.rendezvous(flat, group if group is not None else dist.group.WORLD)
            self._tab_flat = flat
            self.table = flat[: (nu + ni) * embed_dim].view(nu + ni, embed_dim)
            self._shard_ptrs = [int(p) for p in self._tab_hdl.buffer_ptrs]
        else:
            self.table = torch.empty(nu + ni, embed_dim, **f32)
            self._shard_ptrs = [self.table.data_ptr()] if W == 1 else None
        self.user_table, self.item_table = self.table[:nu], self.table[nu:]
        self.err_flag = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self._nu_by_rank = torch.tensor([shard_rows(self.n_user_rows, W, k) for k in range(W)], dtype=torch.int64, device=self.dev)
        if init is not None:
            # parity runs: slice a full (single-process) state dict
            self.user_table.copy_(init["user_tower.embedding.weight"][r::W])
            self.item_table.copy_(init["item_tower.embedding.weight"][r::W])
            flat = lambda t: torch.cat([init[f"{t}_tower.mlp.0.weight"].reshape(-1), init[f"{t}_tower.mlp.0.bias"].reshape(-1),
                                        init[f"{t}_tower.mlp.3.weight"].reshape(-1), init[f"{t}_tower.mlp.3.bias"].reshape(-1)])
            self.user_mlp, self.item_mlp = flat("user").to(**f32).contiguous(), flat("item").to(**f32).contiguous()
        else:
            # counter-based init on the device, same bounds as the reference (xavier-uniform tables incl. row 0,
            # U(±1/√fan_in) Linear layers); MLPs use one seed on every rank (replicated), tables a per-rank seed
            g = torch.Generator(device=self.dev).manual_seed(seed * 1000 + 17 + r)
            a_u, a_i = math.sqrt(6.0 / (self.n_user_rows + embed_dim)), math.sqrt(6.0 / (self.n_item_rows + embed_dim))
            self.user_table.copy_((torch.rand(nu, embed_dim, generator=g, **f32) * 2 - 1) * a_u)
            self.item_table.copy_((torch.rand(ni, embed_dim, generator=g, **f32) * 2 - 1) * a_i)
            g2 = torch.Generator(device=self.dev).manual_seed(seed * 1000 + 3)

            def mlp(din):
                parts = [(torch.rand(hidden_dim * din, generator=g2, **f32) * 2 - 1) / math.sqrt(din),
                         (torch.rand(hidden_dim, generator=g2, **f32) * 2 - 1) / math.sqrt(din),
                         (torch.rand(embed_dim * hidden_dim, generator=g2, **f32) * 2 - 1) / math.sqrt(hidden_dim),
                         (torch.rand(embed_dim, generator=g2, **f32) * 2 - 1) / math.sqrt(hidden_dim)]
                return torch.cat(parts).contiguous()
            self.user_mlp, self.item_mlp = mlp(embed_dim), mlp(embed_dim + n_genres)
        self.state = {k: torch.zeros_like(getattr(self, k)) for k in ("table", "user_mlp", "item_mlp")}
        self.state_v = {k: torch.zeros_like(getattr(self, k)) for k in ("table", "user_mlp", "item_mlp")}
        for d in (self.state, self.state_v):
            d["user_table"], d["item_table"] = d["table"][:nu], d["table"][nu:]
        self.slot = torch.full((max(nu + ni, 1),), -1, dtype=torch.int32, device=self.dev) if adam_mode == "dense" else None
        st = OptState()
        st.lr, st.beta1, st.beta2, st.eps, st.weight_decay, st.max_norm = lr, betas[0], betas[1], eps, weight_decay, max_norm
        st.one_minus_beta1, st.one_minus_beta2, st.beta2_f, st.clip_coef = 1 - betas[0], 1 - betas[1], betas[1], 1.0
        self.opt = torch.zeros(128, dtype=torch.uint8, device=self.dev)
        self._opt_host = st
        self.ops.write_opt(self.opt, st)
        self.steps = 0
        self.max_norm = max_norm
        self.overflow = torch.zeros(1, dtype=torch.int64, device=self.dev)     # padded exchange: requests that did not fit
        self._mlp_sym = None
        if self._symm is not None:
            # p2p exchange: the MLP gradients and the step's scalars are reduced over peer memory too (no NCCL inside the step)
            grp = group if group is not None else dist.group.WORLD
            n_mlp = (self.user_mlp.numel() + self.item_mlp.numel() + 3) // 4 * 4
            gm = self._symm.empty(n_mlp, dtype=torch.float32, device=self.dev)
            sc = self._symm.empty(4, dtype=torch.float32, device=self.dev)
            gm.zero_(); sc.zero_()
            hm, hs = self._symm.rendezvous(gm, grp), self._symm.rendezvous(sc, grp)
            self._mlp_sym = (gm, sc, [int(p) for p in hm.buffer_ptrs], [int(p) for p in hs.buffer_ptrs], (hm, hs))

    # views of the flat MLP blocks
    def _mlp_views(self, flat: torch.Tensor, din: int):
        H, D = self.H, self.D
        o = 0
        W1 = flat[o:o + H * din]; o += H * din
        b1 = flat[o:o + H]; o += H
        W2 = flat[o:o + D * H]; o += D * H
        b2 = flat[o:o + D]
        return W1, b1, W2, b2

    def _route(self, user_ids: torch.Tensor, item_ids: torch.Tensor) -> Route:
        """Requests of BOTH tables in one plan: sample order = [user ids | item ids], bucket order = stable by owner."""
        W = self.world
        n = user_ids.numel() + item_ids.numel()
        perm, inv, local_rows, send = self.ops.route(user_ids.contiguous(), item_ids.contiguous(), W, self._nu_by_rank)
        rt = Route(perm=perm, inv=inv, local_rows=local_rows, send_counts=None)
        if W == 1:
            rt.send_counts = rt.recv_counts = [n]
            rt.recv_rows = rt.local_rows
            return rt
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)
        both = torch.stack([send, recv]).tolist()                                   # the ONE host synchronisation of the step
        rt.send_counts, rt.recv_counts = [int(c) for c in both[0]], [int(c) for c in both[1]]
        rt.recv_rows = all_to_all_var(rt.local_rows, rt.send_counts, rt.recv_counts, self.group)
        return rt

    def capacity(self, n_requests: int) -> int:
        """rows one (source, owner) pair can exchange per step in the padded mode"""
        return min(n_requests, int(math.ceil(n_requests / self.world * self.capacity_factor)) + 16)

    def _route_padded(self, user_ids: torch.Tensor, item_ids: torch.Tensor):
        """Fixed-capacity form of the plan: bucket w of the request list occupies slots [w·C, (w+1)·C) of a padded buffer
        (empty slots carry row -1), so every all-to-all has equal, host-known splits — nothing is read back from the device
        and the whole step can be captured in a CUDA graph.  Requests beyond a bucket's capacity are counted in
        ``self.overflow`` (see ``check_exchange``).  → (slot of each sample-order request [n], rows requested FROM this
        rank [W·C] with -1 padding, C)"""
        W = self.world
        n = user_ids.numel() + item_ids.numel()
        C = self.capacity(n)
        dev = user_ids.device
        if hasattr(self.ops, "route_padded"):                  # the product: one kernel sequence (rb200_route_plan_padded)
            slot_s, send_rows = self.ops.route_padded(user_ids.contiguous(), item_ids.contiguous(), W, self._nu_by_rank, C, self.overflow)
        else:                                                  # the same plan with generic tensor ops (CPU test back end)
            slot_s, send_rows = route_padded_reference(user_ids, item_ids, W, self._nu_by_rank, C, self.overflow)
        if W == 1:
            recv_rows = send_rows[:W * C]
        else:
            recv_rows = torch.empty(W * C, dtype=torch.int64, device=dev)
            dist.all_to_all_single(recv_rows, send_rows[:W * C].contiguous(), group=self.group)
        return slot_s, recv_rows, C

    _side = None

    def _side_stream(self, i: int = 0):
        """the step's side streams (forked from and joined to the step's stream with events: capturable)"""
        if self._side is None:
            # high priority: the exchange plan, the row lists and the sort of the received rows are many small kernels that must not
            # queue behind the towers' CTAs.  (Running the user tower's backward on a second side stream was tried: the towers' phase
            # shrank by 20 us but the replay period did not — C4 at world 1 0.273 -> 0.278 ms.)
            self._side = [torch.cuda.Stream(device=self.dev, priority=-1), torch.cuda.Stream(device=self.dev)]
        return self._side[i]

    def _barrier(self, channel: int) -> None:
        """cross-GPU barrier on the stream (symmetric-memory signal pads; capturable; traps after 60 s instead of hanging)"""
        if self._tab_hdl is not None:
            self._tab_hdl.barrier(channel=channel, timeout_ms=60000)

    def _p2p_buckets(self, n_requests: int):
        """receive buckets of this rank for batches of ``n_requests`` requests: gradient rows [W·C, D] and owner-local rows [W·C],
        bucket k written by rank k — symmetric memory when W > 1 (allocated and exchanged once per batch size, collectively)"""
        Cc = self.capacity(n_requests)
        if Cc not in self._buckets:
            W, D = self.world, self.D
            if self._symm is not None:
                grp = self.group if self.group is not None else dist.group.WORLD
                g = self._symm.empty(W * Cc * D, dtype=torch.float32, device=self.dev)
                r = self._symm.empty(W * Cc, dtype=torch.int64, device=self.dev)
                hg, hr = self._symm.rendezvous(g, grp), self._symm.rendezvous(r, grp)
                self._buckets[Cc] = (g.view(W * Cc, D), r, [int(p) for p in hg.buffer_ptrs], [int(p) for p in hr.buffer_ptrs], (hg, hr))
            else:
                g = torch.empty(W * Cc, D, dtype=torch.float32, device=self.dev)
                r = torch.empty(W * Cc, dtype=torch.int64, device=self.dev)
                self._buckets[Cc] = (g, r, [g.data_ptr()], [r.data_ptr()], None)
        return self._buckets[Cc]

    def check_ids(self) -> None:
        """p2p exchange: raise if an id of a past step was outside its table (synchronises)"""
        if int(self.err_flag.item()) & 1:
            raise IndexError("an id in a previous batch was outside its embedding table")

    def close(self) -> None:
        """Release the captured step.  A CUDA graph that holds NCCL collectives must be gone BEFORE
        ``torch.distributed.destroy_process_group()`` — destroying the communicator under a live graph hangs the process at
        exit (measured on NCCL 2.28 / torch 2.11) — so call this (or drop the trainer) first."""
        self._graph = None
        self._static_loss = None

    def check_exchange(self) -> None:
        """Padded exchange only: raise if any request of a past step did not fit its bucket (synchronises)."""
        lost = int(self.overflow.item())
        if lost:
            raise _lib.RB200Error(f"{lost} embedding-row requests exceeded the exchange capacity (capacity_factor="
                                  f"{self.capacity_factor}): the affected steps are wrong — rebuild the trainer with a larger "
                                  "capacity_factor or exchange='exact'")

    def step(self, user_ids, pos_ids, pos_genres, neg_ids, neg_genres) -> torch.Tensor:
        """One optimiser step on this rank's local batch (device tensors).  Returns the global mean loss (device scalar).

        Exchanges per step: request counts, requested rows, served rows, row gradients (all-to-all), MLP gradients and the
        {Σg², loss} pair (all-reduce).  ``exchange='exact'``: variable-size all-to-alls, one host synchronisation (their split
        sizes).  ``exchange='padded'``: fixed-capacity all-to-alls, no host synchronisation; with ``use_cuda_graph`` the whole
        step, collectives included, is one graph replay after two eager steps."""
        if self.exchange in ("padded", "p2p") and self.check_every > 0 and self.steps > 0 and self.steps % self.check_every == 0:
            self.check_exchange()
        if not self.use_graph:
            return self._step(user_ids, pos_ids, pos_genres, neg_ids, neg_genres)
        batch = (user_ids, pos_ids, pos_genres, neg_ids, neg_genres)
        if self._static_in is None or any(a.shape != b.shape for a, b in zip(self._static_in, batch)):
            self._static_in = [b.clone() for b in batch]
            self._graph, self._eager_steps = None, 0
        else:
            for dst, src in zip(self._static_in, batch):
                dst.copy_(src, non_blocking=True)
        if self._graph is None and self._eager_steps < 2:          # communicators, kernel attributes and allocator warm up
            self._eager_steps += 1
            return self._step(*self._static_in)
        if self._graph is None:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._static_loss = self._step(*self._static_in)
            self._graph = g
            self.steps -= 1                                        # capture enqueues nothing
        self._graph.replay()
        self.steps += 1
        return self._static_loss

    def train_epoch(self, batches) -> float:
        """The reference's ``train_epoch`` loop (``train_embeddings.py:170-199``) over this rank's batches, given as 5-tuples of PINNED
        host tensors ``(user_ids, pos_ids, pos_genres, neg_ids, neg_genres)`` of one shape: batch i+1 travels host → device on a copy
        stream into the other of two staging sets while step i runs, every step's loss goes device → host into a pinned array
        (4 bytes, asynchronous), and the host synchronises ONCE, at the end.  Returns the mean of the steps' global mean losses.
        Collective over the ranks (every rank passes the same number of batches)."""
        n = len(batches)
        if n == 0:
            return 0.0
        main = torch.cuda.current_stream(self.dev)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.dev)
        cs = self._copy_stream
        stage = [[torch.empty(t.shape, dtype=t.dtype, device=self.dev) for t in batches[0]] for _ in range(2)]
        copied = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        losses = torch.empty(n, dtype=torch.float32).pin_memory()

        def upload(i: int) -> None:
            s = i % 2
            with torch.cuda.stream(cs):
                if i >= 2:
                    cs.wait_event(consumed[s])                     # step i − 2 has taken this staging set
                else:
                    cs.wait_stream(main)
                for dst, src in zip(stage[s], batches[i]):
                    dst.copy_(src, non_blocking=True)
                copied[s].record(cs)

        upload(0)
        for i in range(n):
            if i + 1 < n:
                upload(i + 1)
            main.wait_event(copied[i % 2])
            loss = self.step(*stage[i % 2])                        # (copies the staging set into the captured step's inputs, replays)
            consumed[i % 2].record(main)
            losses[i:i + 1].copy_(loss, non_blocking=True)
        torch.cuda.synchronize(self.dev)
        return float(losses.mean())

    _marks = None
    _stamp_buf = None
    _stamp_names: Optional[list] = None

    def _mark(self, name: str) -> None:
        if self._stamp_buf is not None:                            # device timestamps: capturable (profile_stages_graph)
            i = len(self._stamp_names)
            self._stamp_names.append(name)
            check(self.ops.lib.rb200_stamp(self._stamp_buf.data_ptr(), i, stream_ptr()), "stamp")
        elif self._marks is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self._marks.append((name, e))

    def profile_stages_graph(self, user_ids, pos_ids, pos_genres, neg_ids, neg_genres, reps: int = 9) -> Dict[str, float]:
        """Phase times INSIDE a CUDA-graph replay of the step (ms, median of ``reps`` replays): the step is captured once more with
        a one-thread ``%globaltimer`` kernel at every phase boundary (``rb200_stamp``; ≈ 2 us each, which the sum includes).  Unlike
        :meth:`profile_stages` no host launch time is in these numbers.  The replays are real optimiser steps.  Collective over
        the ranks (every rank must call it)."""
        batch = [b.clone() for b in (user_ids, pos_ids, pos_genres, neg_ids, neg_genres)]
        self._stamp_buf = torch.zeros(64, dtype=torch.int64, device=self.dev)
        try:
            self._stamp_names = []
            self._mark("begin")
            self._step(*batch)                                     # eager once: allocator warm-up for the stamped variant
            torch.cuda.synchronize(self.dev)
            self._stamp_names = []
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._mark("begin")
                self._step(*batch)
            self.steps -= 1
            names = list(self._stamp_names)
            acc: Dict[str, list] = {}
            for rep in range(reps + 2):
                g.replay()
                self.steps += 1
                torch.cuda.synchronize(self.dev)
                if rep < 2:
                    continue
                t = self._stamp_buf[:len(names)].tolist()
                for i in range(1, len(names)):
                    acc.setdefault(names[i], []).append((t[i] - t[i - 1]) * 1e-6)
            del g
        finally:
            self._stamp_buf, self._stamp_names = None, None
        return {k: float(sorted(v)[len(v) // 2]) for k, v in acc.items()}

    def profile_stages(self, user_ids, pos_ids, pos_genres, neg_ids, neg_genres, reps: int = 5) -> Dict[str, float]:
        """Time between CUDA events at the phase boundaries of EAGER steps (ms, median of ``reps`` after one untimed step that
        lets the eager allocator pool grow).  The steps are real optimiser steps.  Eager launches include the host's enqueue
        time wherever the device runs ahead of it: read the SHARES; the graph replay is faster than the sum."""
        acc: Dict[str, list] = {}
        for rep in range(reps + 1):
            self._marks = []
            self._mark("begin")
            self._step(user_ids, pos_ids, pos_genres, neg_ids, neg_genres)
            marks, self._marks = self._marks, None
            torch.cuda.synchronize(self.dev)
            if rep == 0:
                continue
            for (_, a), (name, b) in zip(marks[:-1], marks[1:]):
                acc.setdefault(name, []).append(a.elapsed_time(b))
        return {k: float(sorted(v)[len(v) // 2]) for k, v in acc.items()}

    def _step(self, user_ids, pos_ids, pos_genres, neg_ids, neg_genres) -> torch.Tensor:
        ops, D, H, E, W = self.ops, self.D, self.H, self.E, self.world
        B = user_ids.numel()
        dev = user_ids.device
        f32 = dict(dtype=torch.float32, device=dev)
        padded = self.exchange == "padded"
        p2p = self.exchange == "p2p"
        # steps 1-3: route the ids to their owners, gather there, bring the rows back (bucket order)
        if p2p:
            # peer-memory form: the rows are READ from the owners' shards over NVLink, in sample order — one kernel, no collective.
            # The barrier orders this read after every rank's Adam of the previous step.
            item_ids = torch.cat([pos_ids, neg_ids])
            self._barrier(0)
            self._mark("barrier0")
            # Everything that depends only on the ids runs on a side stream under the gather and the towers: the exchange plan, the
            # plan's row lists to the owners (their own barrier channel), and the owners' sort of the rows they are about to receive
            # (first phase of the segment sum).  The gradient rows follow after the towers' backward.
            side = self._side_stream()
            main = torch.cuda.current_stream(dev)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                C = self.capacity(3 * B)
                slot, send_rows = ops.route_padded(user_ids.contiguous(), item_ids, W, self._nu_by_rank, C, self.overflow)
                g_buf, r_buf, g_ptrs, r_ptrs, _ = self._p2p_buckets(3 * B)
                ops.push_row_lists(r_ptrs, self.rank, C, send_rows)
                self._barrier(3)
                rows_sc = r_buf
                if self.rank == 0:       # the global padding ids 0 live on rank 0 (local rows 0 and n_user_local): no gradient
                    nu0 = self.user_table.shape[0]
                    rows_sc = torch.where((rows_sc == 0) | (rows_sc == nu0), torch.full_like(rows_sc, -1), rows_sc)
                uq, n_uq, scat_ws = ops.scatter_plan(rows_sc, max(self.table.shape[0], 1), -1)
            split_fwd = W > 1      # several GPUs: the item rows cross NVLink on a second side stream under the user tower's forward
            if split_fwd:
                rows = torch.empty(3 * B, D, **f32)
                none = user_ids[:0]
                side2 = self._side_stream(1)
                side2.wait_stream(main)
                with torch.cuda.stream(side2):
                    ops.gather_rows_sharded(self._shard_ptrs, self._nu_host, none, item_ids, self.n_user_rows, self.n_item_rows, D,
                                            self.err_flag, out=rows[B:])
                ops.gather_rows_sharded(self._shard_ptrs, self._nu_host, user_ids.contiguous(), none, self.n_user_rows, self.n_item_rows, D,
                                        self.err_flag, out=rows[:B])
            else:
                rows = ops.gather_rows_sharded(self._shard_ptrs, self._nu_host, user_ids.contiguous(), item_ids, self.n_user_rows,
                                               self.n_item_rows, D, self.err_flag)
            if getattr(self, "_ident", None) is None or self._ident.numel() != 3 * B:
                self._ident = torch.arange(3 * B, dtype=torch.int64, device=dev)
            rt = Route(perm=None, inv=self._ident, local_rows=None, send_counts=None, recv_rows=None)
        elif padded:
            slot, recv_rows, C = self._route_padded(user_ids, torch.cat([pos_ids, neg_ids]))
            served = ops.gather_rows(self.table, recv_rows.clamp(min=0))
            # one row more than the exchange carries: slot W·C is the dummy slot of requests that overflowed their bucket — they
            # read a zero row here and their gradient is dropped below, so an overflow costs samples, never another row's data
            rows = torch.empty(W * C + 1, D, **f32)
            rows[W * C:].zero_()
            if W == 1:
                rows[:W * C].copy_(served)
            else:
                dist.all_to_all_single(rows[:W * C], served, group=self.group)
            rt = Route(perm=None, inv=slot, local_rows=None, send_counts=None, recv_rows=recv_rows)
        else:
            rt = self._route(user_ids, torch.cat([pos_ids, neg_ids]))
            served = ops.gather_rows(self.table, rt.recv_rows)
            rows = served if W == 1 else all_to_all_var(served, rt.recv_counts, rt.send_counts, self.group)
        ops.begin_step(self.opt)
        self._mark("route_gather_exchange")

        uW1, ub1, uW2, ub2 = self._mlp_views(self.user_mlp, D)
        iW1, ib1, iW2, ib2 = self._mlp_views(self.item_mlp, D + E)
        out = torch.empty(3 * B, D, **f32)
        hid = torch.empty(3 * B, H, **f32)
        den = torch.empty(3 * B, **f32)
        inv_u, inv_p, inv_n = rt.inv[:B].contiguous(), rt.inv[B:2 * B].contiguous(), rt.inv[2 * B:].contiguous()
        jobs = [
            dict(table=rows, ids=inv_u, extra=None, W1=uW1, b1=ub1, W2=uW2, b2=ub2, out=out[:B], hid=hid[:B], denom=den[:B]),
            dict(table=rows, ids=inv_p, extra=pos_genres, W1=iW1, b1=ib1, W2=iW2, b2=ib2, out=out[B:2 * B], hid=hid[B:2 * B], denom=den[B:2 * B]),
            dict(table=rows, ids=inv_n, extra=neg_genres, W1=iW1, b1=ib1, W2=iW2, b2=ib2, out=out[2 * B:], hid=hid[2 * B:], denom=den[2 * B:]),
        ]
        drop_p = self.dropout
        # masks keyed by (seed of this rank; optimizer step read on the device, so a graph replay draws new ones)
        seed_r, step_dev = self.seed + 7919 * self.rank, self.opt.data_ptr() + OptState.step.offset
        if p2p and split_fwd:
            # user tower as soon as its rows are here, item towers when theirs have arrived (own launch: own mask stream)
            ops.towers_fwd(jobs[:1], D, H, drop_p, seed_r if drop_p > 0.0 else 0, 0, step_dev if drop_p > 0.0 else None)
            main.wait_stream(side2)
            ops.towers_fwd(jobs[1:], D, H, drop_p, seed_r + 1000003 if drop_p > 0.0 else 0, 0, step_dev if drop_p > 0.0 else None)
        elif drop_p > 0.0:
            ops.towers_fwd(jobs, D, H, drop_p, seed_r, 0, step_dev)
        else:
            ops.towers_fwd(jobs, D, H, 0.0, 0, 0)
        self._mark("towers_fwd")
        loss, du, dp, dn = ops.bpr_pair(out[:B], out[B:2 * B], out[2 * B:], 1.0 / W)
        self._mark("loss")

        dpre, dact, drows = torch.empty(3 * B, D, **f32), torch.empty(3 * B, H, **f32), torch.empty(3 * B, D, **f32)
        Pu, Pi = self.user_mlp.numel(), self.item_mlp.numel()
        sym_red = p2p and self._mlp_sym is not None
        g_mlp = self._mlp_sym[0][:Pu + Pi] if sym_red else torch.empty(Pu + Pi, **f32)     # (peer-readable in the p2p exchange)
        for j, dY, sl in ((jobs[0], du, slice(0, B)), (jobs[1], dp, slice(B, 2 * B)), (jobs[2], dn, slice(2 * B, 3 * B))):
            j.update(dY=dY, dpre=dpre[sl], dact=dact[sl], dRows=drows[sl])
        ops.towers_bwd(jobs[:1], D, H, drop_p, g_mlp[:Pu])
        ops.towers_bwd(jobs[1:], D, H, drop_p, g_mlp[Pu:])
        self._mark("towers_bwd")

        # steps 4-5: row gradients (sample order → bucket order) back to the owning shards, deterministic segment sums there
        if p2p:
            # the exchange plan is computed locally (bucket + slot of every request); the gradient rows and the plan's row list are
            # WRITTEN into the owners' receive buckets over NVLink; the barrier makes them visible before the owners' segment sums
            main.wait_stream(side)
            self._mark("route_plan")
            ops.push_rows_sharded(g_ptrs, r_ptrs, self.rank, C, drows, slot, None)
            self._mark("push_rows")
            self._barrier(1)
            g_rows = g_buf
        elif padded:
            g_pad = torch.empty(W * C + 1, D, **f32)               # empty slots are skipped at the owner (their row is -1);
            g_pad.index_copy_(0, rt.inv, drows)                    # row W·C collects the overflowed requests and is not sent
            if W == 1:
                g_rows = g_pad[:W * C]
            else:
                g_rows = torch.empty(W * C, D, **f32)
                dist.all_to_all_single(g_rows, g_pad[:W * C], group=self.group)
        else:
            g_rows = ops.gather_rows(drows, rt.perm)
            if W > 1:
                g_rows = all_to_all_var(g_rows, rt.send_counts, rt.recv_counts, self.group)
        if p2p:
            self._mark("grad_exchange")
            if sym_red:          # the MLP gradients of all ranks are complete (barrier 1): their reduction runs beside the segment sums
                gm, sc, gm_ptrs, sc_ptrs, _ = self._mlp_sym
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    g_red = torch.empty(gm.numel(), **f32)
                    ops.allreduce_oneshot(gm_ptrs, gm.numel(), g_red)
            ug = ops.scatter_apply(g_rows, max(self.table.shape[0], 1), uq, n_uq, scat_ws)
        else:
            rows_sc = rt.recv_rows
            if self.rank == 0:       # the global padding ids 0 live on rank 0 (local rows 0 and n_user_local): no gradient
                nu0 = self.user_table.shape[0]
                rows_sc = torch.where((rows_sc == 0) | (rows_sc == nu0), torch.full_like(rows_sc, -1), rows_sc)
            self._mark("grad_exchange")
            uq, ug, n_uq = ops.scatter_rows(rows_sc, g_rows, max(self.table.shape[0], 1), -1)
        self._mark("scatter")
        opt64, opt32 = self.opt.view(torch.float64), self.opt.view(torch.float32)
        if sym_red:
            # peer-memory reductions (every rank adds all ranks' buffers in rank order: identical everywhere, no broadcast, no NCCL):
            # the MLP gradients were complete before barrier 1; the scalars need one more barrier after the local segment sums
            ops.sumsq(self.opt, [(ug, n_uq, D)])                           # Σg² of this rank's shard rows
            ops.scalars_publish(self.opt, loss, 1.0 / W, sc)
            self._mark("sumsq_publish")
            self._barrier(2)
            self._mark("barrier2")
            main.wait_stream(side)                                         # the reduced MLP gradients
            g_mlp = g_red[:Pu + Pi]
            # opt.sumsq = Σ over shards + the (replicated) MLP gradient counted once, opt.loss = global mean loss, clip coefficient
            ops.scalars_finish(sc_ptrs, g_mlp, self.opt)
        else:
            if W > 1:
                dist.all_reduce(g_mlp, group=self.group)                   # Σ over ranks of (1/W)-scaled local gradients
            # step 6: global gradient norm and mean loss, on the device.  Table shards are disjoint → their Σg² add up; the MLP
            # gradient is replicated → counted once (on rank 0).
            segs = [(ug, n_uq, D)]
            if self.rank == 0:
                segs.append((g_mlp, None, 0))
            ops.sumsq(self.opt, segs)
            if W > 1:
                red = torch.cat([opt64[_OPT_SUMSQ_F64:_OPT_SUMSQ_F64 + 1], loss.to(torch.float64) / W])
                dist.all_reduce(red, group=self.group)
                opt64[_OPT_SUMSQ_F64:_OPT_SUMSQ_F64 + 1].copy_(red[0:1])
                opt32[_OPT_LOSS_F32:_OPT_LOSS_F32 + 1].copy_(red[1:2].to(torch.float32))
            else:
                opt32[_OPT_LOSS_F32:_OPT_LOSS_F32 + 1].copy_(loss)          # one process: the sum of squares is already complete
        if not sym_red:
            ops.grad_norm_clip(self.opt)
        self._mark("allreduce_norm_clip")

        mlp_aligned = Pu % 4 == 0 and all(t.data_ptr() % 16 == 0 for t in (self.user_mlp, self.item_mlp, g_mlp))
        if self.adam_mode == "dense" or not mlp_aligned or not hasattr(ops, "adam_rows_dense2"):     # (stand-in ops of the CPU tests)
            ops.adam_dense(self.user_mlp, g_mlp[:Pu], self.state["user_mlp"], self.state_v["user_mlp"], self.opt)
            ops.adam_dense(self.item_mlp, g_mlp[Pu:], self.state["item_mlp"], self.state_v["item_mlp"], self.opt)
            if self.adam_mode == "dense":
                ops.adam_table_dense(self.table, self.state["table"], self.state_v["table"], uq, ug, n_uq, self.slot, self.opt)
            else:
                ops.adam_rows(self.table, self.state["table"], self.state_v["table"], uq, ug, n_uq, self.opt)
        else:                   # touched rows of the shard + both MLP blocks: one launch
            ops.adam_rows_dense2(self.table, self.state["table"], self.state_v["table"], uq, ug, n_uq,
                                 [(self.user_mlp, g_mlp[:Pu], self.state["user_mlp"], self.state_v["user_mlp"]),
                                  (self.item_mlp, g_mlp[Pu:], self.state["item_mlp"], self.state_v["item_mlp"])], self.opt)
        self._mark("adam")
        self.steps += 1
        return opt32[_OPT_LOSS_F32:_OPT_LOSS_F32 + 1].clone()

    def full_state(self) -> Dict[str, torch.Tensor]:
        """Gather the shards into a single-process ``state_dict`` layout on every rank (tests / checkpoints)."""
        W, D = self.world, self.D
        out = {}
        for name, local, n_rows in (("user", self.user_table, self.n_user_rows), ("item", self.item_table, self.n_item_rows)):
            full = torch.zeros(n_rows, D, dtype=torch.float32, device=self.dev)
            if W == 1:
                full.copy_(local)
            else:
                rows_max = shard_rows(n_rows, W, 0)
                pad = torch.zeros(rows_max, D, dtype=torch.float32, device=self.dev)
                pad[: local.shape[0]] = local
                parts = [torch.empty_like(pad) for _ in range(W)]
                dist.all_gather(parts, pad, group=self.group)
                for r in range(W):
                    full[r::W] = parts[r][: shard_rows(n_rows, W, r)]
            out[f"{name}_tower.embedding.weight"] = full
        for name, flat, din in (("user", self.user_mlp, D), ("item", self.item_mlp, D + self.E)):
            W1, b1, W2, b2 = self._mlp_views(flat, din)
            out[f"{name}_tower.mlp.0.weight"] = W1.view(self.H, din).clone()
            out[f"{name}_tower.mlp.0.bias"] = b1.clone()
            out[f"{name}_tower.mlp.3.weight"] = W2.view(D, self.H).clone()
            out[f"{name}_tower.mlp.3.bias"] = b2.clone()
        return out


# --------------------------------------------------------------------------------------------------------- #
# sharded exhaustive retrieval (C5)
# --------------------------------------------------------------------------------------------------------- #
def sharded_flat_search(queries: torch.Tensor, local_db: torch.Tensor, k: int, id_base: int, group=None, search=None, merge=None
                        ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-shard top-k + all-gather + merge.  ``local_db`` holds rows [id_base, id_base + n_local); every rank passes
    the same queries and gets the same global top-k.  Shards must be ordered by rank (rank r's ids below rank r+1's)
    so that ties resolve to the lower id, as in the unsharded search."""
    from .faiss_index import flat_search, topk_merge
    search = search or flat_search
    merge = merge or topk_merge
    s, i = search(queries, local_db, k, id_base)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return s, i
    ss = [torch.empty_like(s) for _ in range(world)]
    ii = [torch.empty_like(i) for _ in range(world)]
    dist.all_gather(ss, s.contiguous(), group=group)
    dist.all_gather(ii, i.contiguous(), group=group)
    return merge(torch.stack(ss), torch.stack(ii))


# --------------------------------------------------------------------------------------------------------- #
# sharded IVFFlat retrieval (SURVEY.md §8e "IVF (C3 scaled)")
# --------------------------------------------------------------------------------------------------------- #
class ShardedIVFIndex:
    """IVFFlat over a database whose ROWS are sharded across the ranks: every rank holds all centroids and its slice of every
    inverted list (a complete ``FAISSIndex`` over its own rows, built on the shared coarse quantizer).  A query batch is
    replicated; every rank scans its slices of the probed lists, takes its local top-k, the ``(score, item id)[nq, k]`` lists
    are all-gathered over NCCL and merged (``rb200_topk_merge``), so every rank ends with the same global result.

    With the same centroids the union of the shards' candidates of a query is exactly the candidate set of the unsharded
    index (the probed lists depend only on the query and the centroids), hence the result equals the unsharded search except for
    the order of exact-score ties.  Shards must be ordered by rank (rank r's rows before rank r+1's) so that ties resolve to
    the earlier row, as in the unsharded scan.

    ``index_factory`` / ``merge`` are parameters only so that the host logic can run on CPU/gloo in the tests with stand-ins
    built on the oracle; the product uses ``FAISSIndex`` and ``topk_merge`` (C ABI)."""

    def __init__(self, embed_dim: int = 64, n_lists: int = 100, n_probe: int = 10, group=None, index_factory=None, merge=None):
        from .faiss_index import FAISSIndex, topk_merge
        self.embed_dim, self.n_lists, self.n_probe = embed_dim, n_lists, n_probe
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._factory = index_factory or (lambda: FAISSIndex(embed_dim, n_lists, n_probe))
        self._merge = merge or topk_merge
        self.local = None
        self.ntotal = 0

    def build(self, local_embeddings, local_item_ids, centroids=None) -> None:
        """``local_embeddings`` f32 [n_local, D] (host) = this rank's rows; ``centroids`` [n_lists, D] shared by all ranks —
        when omitted, rank 0 trains the coarse quantizer on its own rows (spherical k-means on the device) and broadcasts it."""
        import numpy as np
        x = np.ascontiguousarray(local_embeddings, dtype=np.float32)
        if centroids is None:
            dev = torch.device("cuda", torch.cuda.current_device())
            if self.rank == 0:
                trainer = self._factory()
                trainer.build_ivf_index(x, list(local_item_ids))
                cen = trainer.index.centroids.clone()
                del trainer
            else:
                cen = torch.empty(self.n_lists, self.embed_dim, dtype=torch.float32, device=dev)
            if self.world > 1:
                dist.broadcast(cen, src=0, group=self.group)
            centroids = cen.cpu().numpy()
        self.local = self._factory()
        self.local.build_ivf_index(x, list(local_item_ids), centroids=np.ascontiguousarray(centroids, dtype=np.float32))
        n = torch.tensor([len(local_item_ids)], dtype=torch.int64, device=self._dev())
        if self.world > 1:
            dist.all_reduce(n, group=self.group)
        self.ntotal = int(n.item())

    def _dev(self):
        st = getattr(self.local, "index", None)
        cen = getattr(st, "centroids", None)
        return cen.device if isinstance(cen, torch.Tensor) else torch.device("cpu")

    def set_n_probe(self, n_probe: int) -> None:
        self.n_probe = n_probe
        self.local.set_n_probe(n_probe)

    def search_device(self, queries: torch.Tensor, k: int = 500) -> Tuple[torch.Tensor, torch.Tensor]:
        """queries: L2-normalised f32 [nq, D] on this rank's device, the same on every rank → (scores [nq, k], item ids [nq, k]),
        -FLT_MAX / -1 padded, identical on every rank."""
        st = self.local.index
        kk = min(k, max(st.ntotal, 1))
        s, i = st.search_device(queries, kk, id_table=self.local._list_item_ids)
        st.check_last_search()
        if kk < k:                                               # a shard with fewer than k rows: pad to the common width
            pad_s = torch.full((s.shape[0], k - kk), -3.4028234663852886e38, dtype=s.dtype, device=s.device)
            pad_i = torch.full((s.shape[0], k - kk), -1, dtype=i.dtype, device=i.device)
            s, i = torch.cat([s, pad_s], 1), torch.cat([i, pad_i], 1)
        if self.world == 1:
            return s, i
        ss = [torch.empty_like(s) for _ in range(self.world)]
        ii = [torch.empty_like(i) for _ in range(self.world)]
        dist.all_gather(ss, s.contiguous(), group=self.group)
        dist.all_gather(ii, i.contiguous(), group=self.group)
        return self._merge(torch.stack(ss), torch.stack(ii))

    def batch_search(self, query_vectors, k: int = 500):
        """host queries (any norm) → host (scores, item ids): ``FAISSIndex.batch_search`` semantics over the whole database"""
        import numpy as np
        from .faiss_index import _normalize_device
        dev = self._dev()
        q = _normalize_device(torch.as_tensor(np.ascontiguousarray(query_vectors, dtype=np.float32), device=dev))
        s, i = self.search_device(q, min(k, self.ntotal))
        return s.cpu().numpy(), i.cpu().numpy()
