"""B200 drop-in for the reference's ``src/models/two_tower.py``.

Same public surface (``UserTower``, ``ItemTower``, ``TwoTowerModel`` with ``forward`` /
``user_tower`` / ``item_tower`` / ``bpr_loss`` / ``in_batch_bpr_loss`` / inference helpers /
``save`` / ``load``), same ``state_dict`` keys, same parameter initialisation, so that the reference's
``train_embeddings.py:153-192``, ``build_index.py:86-105`` and ``serving/recommender.py:148-203`` run on
it unchanged.  The arithmetic, however, is not PyTorch's: every forward and backward below is a call
into ``librb200.so`` (hand-written sm_100a kernels, see ``csrc/``):

===========================  =============================================  ======================
reference                    replaced by                                    kernel source
===========================  =============================================  ======================
two_tower.py:39-42, 68-72    ``rb200_tower_fwd`` (gather+MLP+normalise)     csrc/tower.cu
autograd of the towers       ``rb200_tower_bwd`` + ``rb200_scatter_rows``   csrc/tower.cu, scatter_adam.cu
two_tower.py:117-130         ``rb200_bpr_pair``                             csrc/loss.cu
two_tower.py:132-160         ``rb200_bpr_inbatch``                          csrc/loss.cu, inbatch_tc.cu
===========================  =============================================  ======================

There is no CPU path: tensors must live on a CUDA device or the call raises ``RB200Error``.
"""
from __future__ import annotations

import logging
import os
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import RB200Error, TowerBwdJob, TowerJob, check, ptr, require_cuda, stream_ptr, workspace

logger = logging.getLogger(__name__)

N_GENRES = 18  # MovieLens-1M genre multi-hot width (reference two_tower.py:16)

#: precision mode of the tower MLP kernels: 0 = fp32 FFMA, 1 = tcgen05 TF32, 2 = tcgen05 3xTF32 (fp32-grade).
#: "auto" (default) = 2 where the tensor-core kernels cover the widths (D in {64, 128}, H=128, <= 24 extra columns), else 0.
TOWER_MODE = os.environ.get("RB200_TOWER_MODE", "auto")


def tower_mode_for(D: int, H: int, E: int, mode=None) -> int:
    mode = TOWER_MODE if mode is None else mode
    if str(mode) == "auto":
        return 2 if (D in (64, 128) and H == 128 and E <= 24) else 0
    return int(mode)


#: in-batch BPR kernel: 0 = fp32 FFMA (SIMT), 1 = tcgen05 TF32 (stated fast mode), 2 = tcgen05 3xTF32 (fp32-grade);
#: "auto" = 2 where the tensor-core kernel is built (D = 64, B >= 2), else 0
INBATCH_MODE = os.environ.get("RB200_INBATCH_MODE", "auto")


def inbatch_mode_for(B: int, D: int, mode=None) -> int:
    mode = INBATCH_MODE if mode is None else mode
    if str(mode) == "auto":
        return 2 if (D == 64 and B >= 2) else 0
    return int(mode)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _raise_on_bad_ids(err: torch.Tensor, n_rows: int) -> None:
    flag = int(err.item())
    if flag & 1:
        raise IndexError(f"index out of range in self (an id was outside the embedding table [0, {n_rows - 1}])")
    if flag & 2:
        raise RB200Error("tower kernel: a tensor-core pipeline wait timed out (results of this call are invalid)")


class _TowerFn(torch.autograd.Function):
    """gather → Linear → ReLU → dropout → Linear → L2-normalise, fused (two_tower.py:39-42 / 68-72)."""

    @staticmethod
    def forward(ctx, ids, extra, table, W1, b1, W2, b2, drop_p, seed, offset, keep_mask, mode=None):
        lib = _lib.load()
        dev = require_cuda(ids, extra, table, W1, b1, W2, b2, keep_mask)
        if ids.dtype != torch.int64:
            ids = ids.long()
        shape = ids.shape
        ids = ids.reshape(-1).contiguous()
        B, (n_rows, D), H = ids.numel(), table.shape, W1.shape[0]
        E = 0 if extra is None else extra.shape[-1]
        if extra is not None:
            extra = _f32c(extra.reshape(B, E))
        if W1.shape[1] != D + E:
            raise RB200Error(f"first Linear expects {W1.shape[1]} inputs but embedding({D}) + extra({E}) were given")
        table, W1, b1, W2, b2 = (_f32c(t) for t in (table, W1, b1, W2, b2))
        need_grad = any(ctx.needs_input_grad)
        out = torch.empty(B, D, dtype=torch.float32, device=dev)
        hid = torch.empty(B, H, dtype=torch.float32, device=dev) if need_grad else None
        denom = torch.empty(B, dtype=torch.float32, device=dev) if need_grad else None
        if keep_mask is not None:
            keep_mask = keep_mask.reshape(B, H).to(torch.uint8).contiguous()
        mode = tower_mode_for(D, H, E, mode)
        # an id outside the table is clamped to row 0 by the kernels and reported through this flag; nn.Embedding raises
        # IndexError there, and serving/recommender.py:200-207 relies on the exception for its cold-start fallback
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        if B > 0:
            job = TowerJob(ptr(table), ptr(ids), ptr(extra), ptr(W1), ptr(b1), ptr(W2), ptr(b2), ptr(out), ptr(hid),
                           ptr(denom), ptr(keep_mask), n_rows, B, E, 0, None)
            wsb = lib.rb200_tower_fwd_workspace_bytes(1, D, H, E, mode)
            ws = workspace(wsb, dev) if wsb else None
            with torch.cuda.device(dev):
                check(lib.rb200_tower_fwd(job, 1, D, H, float(drop_p), int(seed), int(offset), None, mode, ptr(err), ptr(ws), wsb,
                                          stream_ptr()), "rb200_tower_fwd")
        if need_grad:
            # training: the flag is read in backward (the caller's loss.item() synchronises every step anyway), so the forward
            # of the three towers stays asynchronous
            ctx.save_for_backward(ids, extra, table, W1, W2, out, hid, denom, err)
            ctx.drop_p = float(drop_p)
            ctx.mode = mode
            ctx.dims = (B, D, H, E, n_rows)
        elif B > 0 and not torch.cuda.is_current_stream_capturing():
            _raise_on_bad_ids(err, n_rows)        # inference: synchronous, like the .cpu() that follows in every caller
        return out.view(*shape, D)

    @staticmethod
    def backward(ctx, dY):
        lib = _lib.load()
        ids, extra, table, W1, W2, out, hid, denom, err = ctx.saved_tensors
        B, D, H, E, n_rows = ctx.dims
        dev = table.device
        if not torch.cuda.is_current_stream_capturing():
            _raise_on_bad_ids(err, n_rows)
        Din = D + E
        P = H * Din + H + D * H + D
        grads = torch.empty(P, dtype=torch.float32, device=dev)
        g_table = torch.zeros(n_rows, D, dtype=torch.float32, device=dev) if ctx.needs_input_grad[2] else None
        if B > 0:
            dY = _f32c(dY.reshape(B, D))
            dpre = torch.empty(B, D, dtype=torch.float32, device=dev)
            dact = torch.empty(B, H, dtype=torch.float32, device=dev)
            drows = torch.empty(B, D, dtype=torch.float32, device=dev)
            wsb = lib.rb200_tower_bwd_workspace_bytes(D, H, E)
            ws = workspace(wsb, dev)
            job = TowerBwdJob(ptr(table), ptr(ids), ptr(extra), n_rows, B, E, 0, ptr(W1), ptr(W2), ptr(dY), ptr(out),
                              ptr(denom), ptr(hid), ptr(dpre), ptr(dact), ptr(drows), None)
            with torch.cuda.device(dev):
                check(lib.rb200_tower_bwd(job, 1, D, H, ctx.drop_p, ctx.mode, ptr(grads), 0, ptr(ws), wsb, stream_ptr()),
                      "rb200_tower_bwd")
                if g_table is not None:
                    # dense nn.Embedding gradient (padding_idx = 0 receives nothing), deterministic
                    sb = lib.rb200_scatter_workspace_bytes(B, n_rows)
                    sws = workspace(sb, dev)
                    check(lib.rb200_scatter_rows(ptr(ids), ptr(drows), B, D, n_rows, 0, ptr(g_table), None, None, None,
                                                 None, ptr(sws), sb, stream_ptr()), "rb200_scatter_rows")
        else:
            grads.zero_()
        o = 0
        gW1 = grads[o:o + H * Din].view(H, Din); o += H * Din
        gb1 = grads[o:o + H]; o += H
        gW2 = grads[o:o + D * H].view(D, H); o += D * H
        gb2 = grads[o:o + D]
        return None, None, g_table, gW1, gb1, gW2, gb2, None, None, None, None, None


class _BprPairFn(torch.autograd.Function):
    """-logsigmoid(u·p − u·n).mean() (two_tower.py:127-129); forward also produces the gradients."""

    @staticmethod
    def forward(ctx, u, p, n):
        lib = _lib.load()
        dev = require_cuda(u, p, n)
        u, p, n = _f32c(u), _f32c(p), _f32c(n)
        B, D = u.shape
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        need = any(ctx.needs_input_grad)
        du, dp, dn = (torch.empty_like(u) for _ in range(3)) if need else (None, None, None)
        wsb = lib.rb200_bpr_pair_workspace_bytes(B)
        ws = workspace(wsb, dev)
        with torch.cuda.device(dev):
            check(lib.rb200_bpr_pair(ptr(u), ptr(p), ptr(n), B, D, ptr(loss), ptr(du), ptr(dp), ptr(dn), 1.0, ptr(ws), wsb,
                                     stream_ptr()), "rb200_bpr_pair")
        if need:
            ctx.save_for_backward(du, dp, dn)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        du, dp, dn = ctx.saved_tensors
        return du * g, dp * g, dn * g


class _BprInBatchFn(torch.autograd.Function):
    """in_batch_bpr_loss (two_tower.py:132-160) without materialising the B×B score matrix."""

    @staticmethod
    def forward(ctx, u, i, mode):
        lib = _lib.load()
        dev = require_cuda(u, i)
        u, i = _f32c(u), _f32c(i)
        B, D = u.shape
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        need = any(ctx.needs_input_grad)
        du, di = (torch.empty_like(u), torch.empty_like(i)) if need else (None, None)
        wsb = lib.rb200_bpr_inbatch_workspace_bytes(B, D)
        ws = workspace(wsb, dev)
        with torch.cuda.device(dev):
            check(lib.rb200_bpr_inbatch(ptr(u), ptr(i), B, D, int(mode), ptr(loss), ptr(du), ptr(di), 1.0, ptr(ws), wsb,
                                        stream_ptr()), "rb200_bpr_inbatch")
        if need:
            ctx.save_for_backward(du, di)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        du, di = ctx.saved_tensors
        return du * g, di * g, None


class _Tower(nn.Module):
    """Parameter container with the reference's module tree (``embedding``, ``mlp.0``, ``mlp.3``) so that
    ``state_dict`` keys, default initialisation and ``.parameters()`` order are those of the reference
    (two_tower.py:25-37, 52-66).  ``mlp`` is never *called*: the fused kernel does the arithmetic."""

    def __init__(self, n_rows: int, embed_dim: int, hidden_dim: int, dropout: float, extra_dim: int):
        super().__init__()
        self.embedding = nn.Embedding(n_rows + 1, embed_dim, padding_idx=0)
        self.mlp = nn.Sequential(
            nn.Linear(embed_dim + extra_dim, hidden_dim),
            nn.ReLU(),
            nn.Dropout(dropout),
            nn.Linear(hidden_dim, embed_dim),
        )
        self.extra_dim = extra_dim
        self.mode = None          # None → module default (RB200_TOWER_MODE / auto); 0, 1 or 2 to force
        self._calls = 0
        # dropout stream of this tower: Philox offset = stream id · 2^40 + call counter, seed = torch.initial_seed() — a fixed
        # torch.manual_seed gives the same masks in every process (user tower 0, item tower 1; the reference is reproducible too)
        self._stream_id = 1 if extra_dim else 0
        self._init_weights()

    def _init_weights(self):
        # xavier over the whole table, including the padding row
        nn.init.xavier_uniform_(self.embedding.weight)

    @property
    def dropout_p(self) -> float:
        return float(self.mlp[2].p)

    def _run(self, ids, extra, keep_mask=None):
        p = self.dropout_p if self.training else 0.0
        self._calls += 1
        seed = torch.initial_seed() & 0x7FFFFFFFFFFFFFFF
        l1, l2 = self.mlp[0], self.mlp[3]
        return _TowerFn.apply(ids, extra, self.embedding.weight, l1.weight, l1.bias, l2.weight, l2.bias, p, seed,
                              (self._stream_id << 40) + self._calls, keep_mask, self.mode)


class UserTower(_Tower):
    """user_id → embedding → MLP → L2-normalised vector."""

    def __init__(self, n_users: int, embed_dim: int, hidden_dim: int = 128, dropout: float = 0.1):
        super().__init__(n_users, embed_dim, hidden_dim, dropout, 0)

    def forward(self, user_ids: torch.Tensor, keep_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self._run(user_ids, None, keep_mask)


class ItemTower(_Tower):
    """(item_id, 18-dim genre multi-hot) → [embedding ; genres] → MLP → L2-normalised vector."""

    def __init__(self, n_items: int, embed_dim: int, hidden_dim: int = 128, dropout: float = 0.1):
        super().__init__(n_items, embed_dim, hidden_dim, dropout, N_GENRES)

    def forward(self, item_ids: torch.Tensor, genre_vectors: torch.Tensor,
                keep_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self._run(item_ids, genre_vectors, keep_mask)


class TwoTowerModel(nn.Module):
    """Drop-in for ``src.models.two_tower.TwoTowerModel`` (reference two_tower.py:75-251)."""

    def __init__(self, n_users: int, n_items: int, embed_dim: int = 64, hidden_dim: int = 128, dropout: float = 0.1):
        super().__init__()
        self.n_users = int(n_users)
        self.n_items = int(n_items)
        self.embed_dim = int(embed_dim)
        self.hidden_dim = int(hidden_dim)
        self.user_tower = UserTower(self.n_users, embed_dim, hidden_dim, dropout)
        self.item_tower = ItemTower(self.n_items, embed_dim, hidden_dim, dropout)
        self._item_embeddings: Optional[torch.Tensor] = None
        self._item_id_to_idx: Optional[Dict[int, int]] = None
        self._idx_to_item_id: Optional[Dict[int, int]] = None

    # -- training surface ------------------------------------------------------------------ #
    def forward(self, user_ids, pos_item_ids, pos_genre_vectors, neg_item_ids=None, neg_genre_vectors=None
                ) -> Tuple[torch.Tensor, torch.Tensor]:
        # like the reference (two_tower.py:107-115) the negative arguments are accepted and ignored
        return self.user_tower(user_ids), self.item_tower(pos_item_ids, pos_genre_vectors)

    def bpr_loss(self, user_emb, pos_item_emb, neg_item_emb) -> torch.Tensor:
        return _BprPairFn.apply(user_emb, pos_item_emb, neg_item_emb)

    def in_batch_bpr_loss(self, user_emb, item_emb, mode: Optional[int] = None) -> torch.Tensor:
        return _BprInBatchFn.apply(user_emb, item_emb, inbatch_mode_for(user_emb.shape[0], user_emb.shape[1], mode))

    # -- inference helpers (two_tower.py:166-210) ------------------------------------------ #
    def _device(self) -> torch.device:
        """The COMPUTE device.  The helpers below take and return host NumPy data, so the `device` argument the reference's
        callers pass (run_pipeline.py:145-177 hardcodes torch.device('cpu')) only says where the caller's tensors live; the
        arithmetic has no CPU path.  Weights still on the CPU are moved to the current CUDA device on first use."""
        dev = self.user_tower.embedding.weight.device
        if dev.type != "cuda" and torch.cuda.is_available():
            dev = torch.device("cuda", torch.cuda.current_device())
            self.to(dev)
        return dev

    @torch.no_grad()
    def get_user_embedding(self, user_id: int, device: torch.device = None) -> np.ndarray:
        self.eval()   # the reference leaves the model in eval mode too
        if not 0 <= int(user_id) <= self.n_users:     # nn.Embedding raises here; recommender.py:200-207 turns it into the fallback
            raise IndexError(f"index out of range in self (user id {user_id}, table [0, {self.n_users}])")
        ids = torch.tensor([int(user_id)], dtype=torch.long, device=self._device())
        return self.user_tower(ids).cpu().numpy()[0]

    @torch.no_grad()
    def get_item_embeddings(self, item_ids: List[int], genre_vectors: np.ndarray, device: torch.device = None,
                            batch_size: int = 512) -> np.ndarray:
        """Embeddings for a catalog slice.  ``batch_size`` is accepted for signature compatibility; rows are
        independent, so the whole list goes through one kernel launch (one H2D, one D2H copy)."""
        self.eval()
        dev = self._device()
        if len(item_ids) == 0:
            return np.zeros((0, self.embed_dim), dtype=np.float32)
        ids = torch.as_tensor(np.asarray(item_ids, dtype=np.int64), device=dev)
        genres = torch.as_tensor(np.ascontiguousarray(genre_vectors, dtype=np.float32), device=dev)
        return self.item_tower(ids, genres).cpu().numpy()

    def precompute_item_embeddings(self, item_ids: List[int], genre_vectors: np.ndarray, device: torch.device = None) -> None:
        embs = self.get_item_embeddings(item_ids, genre_vectors, device)
        self._item_embeddings = torch.tensor(embs, dtype=torch.float32)
        self._item_id_to_idx = {int(iid): idx for idx, iid in enumerate(item_ids)}     # plain ints: loadable with weights_only
        self._idx_to_item_id = {idx: int(iid) for idx, iid in enumerate(item_ids)}
        logger.info("Precomputed %d item embeddings (dim=%d)", len(item_ids), embs.shape[1])

    # -- persistence (two_tower.py:216-251) -------------------------------------------------- #
    def save(self, path: str) -> None:
        """Same checkpoint dict as the reference, so ``.pt`` files interoperate both ways.  ``hidden_dim`` is an
        extra key the reference's ``load`` ignores (it silently assumes 128 — SURVEY.md F8)."""
        save_path = Path(path)
        save_path.parent.mkdir(parents=True, exist_ok=True)
        torch.save({
            "state_dict": {k: v.detach().cpu() for k, v in self.state_dict().items()},
            "n_users": self.n_users,
            "n_items": self.n_items,
            "embed_dim": self.embed_dim,
            "item_id_to_idx": self._item_id_to_idx,
            "idx_to_item_id": self._idx_to_item_id,
            "hidden_dim": self.hidden_dim,
        }, save_path)
        logger.info("Saved two-tower model to %s", save_path)

    @classmethod
    def load(cls, path: str, device: torch.device = torch.device("cpu")) -> "TwoTowerModel":
        # tensors + int dicts; no arbitrary pickles.  Checkpoints written by the reference hold NumPy integer scalars
        # (`n_users = ratings_df["user_id"].max()`, train_embeddings.py:135-136), which the safe unpickler admits by name.
        import numpy
        scalars = [numpy.dtype, numpy.dtypes.Int64DType, numpy.dtypes.Int32DType, numpy.dtypes.Float64DType, numpy.dtypes.Float32DType]
        try:
            scalars.append(numpy._core.multiarray.scalar)
        except AttributeError:                                       # numpy < 2
            scalars.append(numpy.core.multiarray.scalar)
        with torch.serialization.safe_globals(scalars):
            ck = torch.load(path, map_location="cpu", weights_only=True)
        sd = ck["state_dict"]
        hidden = ck.get("hidden_dim", sd["user_tower.mlp.0.weight"].shape[0])
        model = cls(n_users=int(ck["n_users"]), n_items=int(ck["n_items"]), embed_dim=int(ck["embed_dim"]), hidden_dim=int(hidden))
        model.load_state_dict(sd)
        model._item_id_to_idx = ck.get("item_id_to_idx")
        model._idx_to_item_id = ck.get("idx_to_item_id")
        # `device` is the caller's host-facing device (the reference's default and run_pipeline.py:146 pass cpu); the weights go
        # to the compute device: the requested CUDA device, else the current one when a GPU is present
        device = torch.device(device)
        if device.type != "cuda" and torch.cuda.is_available():
            device = torch.device("cuda", torch.cuda.current_device())
        model.to(device)
        model.eval()
        logger.info("Loaded two-tower model from %s (users=%d, items=%d, dim=%d)", path, model.n_users, model.n_items,
                    model.embed_dim)
        return model
