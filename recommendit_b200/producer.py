"""Device-side batch producer — replaces ``UserItemDataset`` + ``DataLoader`` of the reference
(src/training/train_embeddings.py:23-79, 144-151) for the fused trainer.

The reference builds each sample in Python (one ``np.random.choice`` per negative) and feeds ~22 k samples/s; the fused step
consumes tens of millions per second.  Here the sample stream is produced by one kernel launch per batch
(``rb200_sample_batch``, csrc/sampler.cu) straight into the trainer's staging buffers:

    producer = DeviceBatchProducer(ratings_user_ids, ratings_item_ids, ratings, all_item_ids, n_users, seed=0)
    trainer = FusedBPRTrainer(model, item_extra_table=genre_table)          # genres are looked up by item id in the kernels
    for epoch in range(epochs):
        mean_loss = producer.train_epoch(trainer, batch_size, epoch)

Same semantics as the reference: positives are the pairs with rating >= ``min_rating``, every epoch visits them in a fresh
pseudo-random order in full batches (``shuffle=True, drop_last=True``), each sample gets one negative drawn uniformly from the
catalog and redrawn while the user has rated it.  Counter-based (Feistel permutation + Philox): batch (epoch, step) is a pure
function of the seed (the test suite holds a CPU restatement that the kernel must match bit for bit).
"""
from __future__ import annotations

import numpy as np
import torch

import ctypes as C

from . import _lib
from ._lib import RB200Error, Sampler, check, ptr, stream_ptr

_BITMAP_MAX_BYTES = 1 << 28      # rated-items bitmap (one bit per (user, item id)) is built when it fits this; else the CSR is searched


def build_host_index(user_ids, item_ids, ratings, all_item_ids, n_users: int, min_rating: float = 4.0,
                     bitmap_max_bytes: int = 1 << 28) -> dict:
    """One-time index build on the host — what ``UserItemDataset.__init__`` does with pandas (train_embeddings.py:39-48):
    positives = the (user, item) rows with rating >= ``min_rating`` in input order (repeated rows stay), rated = the SET of
    (user, item) pairs of any rating as a CSR (items ascending within a user), plus, when it fits ``bitmap_max_bytes``, the same
    relation as one bit per (user, item id).  Pure NumPy (checked on CPU against the reference's own dataset class:
    tests/golden/sampler.npz)."""
    u = np.asarray(user_ids, dtype=np.int64)
    i = np.asarray(item_ids, dtype=np.int64)
    r = np.asarray(ratings, dtype=np.float64)
    catalog = np.asarray(all_item_ids, dtype=np.int64)
    if u.size == 0 or u.min() < 0 or u.max() > n_users:
        raise ValueError("user ids must lie in [0, n_users]")
    pos = r >= min_rating
    if catalog.size < 1 or int(pos.sum()) < 1:
        raise ValueError("empty catalog or no positive pairs")
    pairs = np.unique(np.stack([u, i], 1), axis=0)
    counts = np.bincount(pairs[:, 0], minlength=n_users + 1)
    offsets = np.zeros(n_users + 2, dtype=np.int64)
    offsets[1:] = np.cumsum(counts)
    out = {"pos_users": u[pos], "pos_items": i[pos], "rated_offsets": offsets, "rated_items": pairs[:, 1].copy(), "catalog": catalog,
           "bitmap": None, "bitmap_words": 0}
    # optional bitmap of the same relation: one 32-bit load instead of a binary search per rejection test (same results)
    words = (max(int(pairs[:, 1].max()), int(catalog.max())) >> 5) + 1
    if (n_users + 1) * words * 4 <= bitmap_max_bytes:
        bm = np.zeros((n_users + 1) * words, dtype=np.uint32)
        np.bitwise_or.at(bm, pairs[:, 0] * words + (pairs[:, 1] >> 5), (np.uint32(1) << (pairs[:, 1] & 31).astype(np.uint32)))
        out["bitmap"], out["bitmap_words"] = bm, words
    return out


class DeviceBatchProducer:
    def __init__(self, user_ids, item_ids, ratings, all_item_ids, n_users: int, min_rating: float = 4.0, seed: int = 0,
                 device=None, rank: int = 0, world: int = 1):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RB200Error("DeviceBatchProducer needs a CUDA device (there is no CPU fallback)")
        self.dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        # data-parallel ranks share one epoch: every step consumes world·B positives, rank r takes the r-th slice of B
        # (same seed on every rank → disjoint batches; DataParallelBPRTrainer averages their gradients)
        if not (world >= 1 and 0 <= rank < world):
            raise ValueError("rank must lie in [0, world)")
        self.rank, self.world = int(rank), int(world)
        ix = build_host_index(user_ids, item_ids, ratings, all_item_ids, n_users, min_rating, _BITMAP_MAX_BYTES)
        self.n_pos = len(ix["pos_users"])
        to = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.int64), device=self.dev)
        self.pos_users, self.pos_items = to(ix["pos_users"]), to(ix["pos_items"])
        self.rated_offsets, self.rated_items = to(ix["rated_offsets"]), to(ix["rated_items"])
        self.catalog = to(ix["catalog"])
        self.rated_bitmap, self.bitmap_words = None, 0
        if ix["bitmap"] is not None:
            self.rated_bitmap = torch.as_tensor(ix["bitmap"].view(np.int32), device=self.dev)
            self.bitmap_words = ix["bitmap_words"]
        self._samplers = {}

    def sampler(self, batch_size: int) -> Sampler:
        """The ``rb200_sampler`` description of this stream at ``batch_size`` (what ``rb200_step_params.next_batch`` points at)."""
        if batch_size not in self._samplers:
            nb = self.batches_per_epoch(batch_size)
            if nb < 1:
                raise ValueError(f"batch_size {batch_size} exceeds the {self.n_pos} positive pairs (drop_last leaves no batch)")
            s = Sampler()
            s.pos_users, s.pos_items, s.n_pos = ptr(self.pos_users), ptr(self.pos_items), self.n_pos
            s.rated_offsets, s.rated_items = ptr(self.rated_offsets), ptr(self.rated_items)
            s.rated_bitmap = ptr(self.rated_bitmap) if self.rated_bitmap is not None else None
            s.bitmap_words = self.bitmap_words
            s.catalog, s.n_cat = ptr(self.catalog), self.catalog.numel()
            s.seed, s.batches_per_epoch = self.seed, nb
            s.rank, s.world = self.rank, self.world
            self._samplers[batch_size] = s
        return self._samplers[batch_size]

    def batches_per_epoch(self, batch_size: int) -> int:
        return self.n_pos // (batch_size * self.world)                      # drop_last=True; a step takes world·batch_size positives

    def fill(self, out_users: torch.Tensor, out_pos: torch.Tensor, out_neg: torch.Tensor, epoch: int, step: int) -> None:
        """Write batch ``step`` of ``epoch`` into three int64 device tensors of the same length (asynchronous)."""
        B = out_users.numel()
        with torch.cuda.device(self.dev):
            check(self.lib.rb200_sample_batch(ptr(self.pos_users), ptr(self.pos_items), self.n_pos, ptr(self.rated_offsets),
                                              ptr(self.rated_items), ptr(self.catalog), self.catalog.numel(), B, self.seed, int(epoch),
                                              int(step), self.rank, self.world, ptr(out_users), ptr(out_pos), ptr(out_neg),
                                              stream_ptr()),
                  "rb200_sample_batch")

    def fill_from_counter(self, batch_size: int, counter: torch.Tensor, out_users, out_pos, out_neg) -> None:
        """Batch number ``counter[0]`` (device int64; epoch = g // batches_per_epoch, step = g % batches_per_epoch) — the form the
        fused step runs by itself at the end of every step."""
        with torch.cuda.device(self.dev):
            check(self.lib.rb200_sample_batch_dev(C.byref(self.sampler(batch_size)), batch_size, ptr(counter), ptr(out_users),
                                                  ptr(out_pos), ptr(out_neg), stream_ptr()), "rb200_sample_batch_dev")

    def train_epoch(self, trainer, batch_size: int, epoch: int, in_graph: bool = True) -> float:
        """One epoch of the reference's training loop (train_embeddings.py:176-199) with the batches produced on the device; the
        per-step losses are summed on the device and read once.  ``trainer`` must have been built with ``item_extra_table``
        (genres by item id).

        ``in_graph=True`` (default): the fused step samples its own next batch on a side stream under the optimizer kernels
        (``rb200_step_params.next_batch``), so an epoch is ``batches_per_epoch`` graph replays and nothing else.  The position
        in the stream is the optimizer's device-resident step counter, so this form needs the trainer to stand exactly at
        the start of ``epoch`` (``epoch * batches_per_epoch`` steps done); otherwise, and with ``in_graph=False``, every
        step gets its own sampling launch (``rb200_sample_batch``).  Both forms produce the same batches bit for bit."""
        if trainer.item_extra_table is None and trainer.E:
            raise RB200Error("DeviceBatchProducer.train_epoch needs a trainer built with item_extra_table (genres by item id)")
        if trainer._B != batch_size:
            trainer._alloc(batch_size)
        nb = self.batches_per_epoch(batch_size)
        total = torch.zeros(1, dtype=torch.float32, device=self.dev)
        in_graph = in_graph and trainer.loss_kind == 0 and trainer._steps_done == epoch * nb and nb > 0
        with torch.cuda.device(self.dev):
            if in_graph:
                if trainer._sampler is None or trainer._sampler[1] is not self or trainer._sampler[0] is not self.sampler(batch_size):
                    trainer.attach_producer(self.sampler(batch_size), self)
                # first batch of the epoch: the same kernel, read from the same counter (opt.step sits at byte 32 of the state)
                counter = trainer.opt_dev[32:40].view(torch.int64)
                self.fill_from_counter(batch_size, counter, trainer.user_ids, trainer.pos_ids, trainer.neg_ids)
                for _ in range(nb):
                    total.add_(trainer.step())
            else:
                if trainer._sampler is not None:
                    trainer.attach_producer(None)
                for step in range(nb):
                    self.fill(trainer.user_ids, trainer.pos_ids, trainer.neg_ids, epoch, step)
                    total.add_(trainer.step())
            return float(total.item()) / nb if nb else float("nan")
