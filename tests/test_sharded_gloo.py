"""World-size-2 CPU (gloo) tests of the multi-GPU host logic: row-sharded training step and sharded retrieval give the
same result as the single-process computation ("results independent of world size").
Arithmetic is supplied by the NumPy oracle stand-in (tests/oracle_ops.py); the exchange plan is the product's."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ivf_oracle as V
from oracle import two_tower_oracle as O

NU, NI, D, H, B = 53, 41, 32, 64, 24


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _batches(rank, step):
    rng = np.random.default_rng(100 * step + rank)
    u, p, n = rng.integers(0, NU + 1, B), rng.integers(0, NI + 1, B), rng.integers(0, NI + 1, B)
    pg, ng = (rng.random((B, 18)) < 0.2).astype(np.float32), (rng.random((B, 18)) < 0.2).astype(np.float32)
    return u, p, pg, n, ng


def _train_worker(rank, world, port, adam_mode, exchange, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from recommendit_b200.sharded import ShardedBPRTrainer
        from tests.oracle_ops import OracleOps
        P = O.init_params(NU, NI, D, H, seed=3)
        init = {k: torch.from_numpy(v) for k, v in P.items()}
        tr = ShardedBPRTrainer(NU, NI, D, H, adam_mode=adam_mode, device="cpu", ops=OracleOps(), init=init, lr=1e-2, exchange=exchange,
                               capacity_factor=1.5)
        losses = []
        for step in range(2):
            b = _batches(rank, step)
            losses.append(float(tr.step(*[torch.from_numpy(a) for a in b])))
        full = {k: v.numpy() for k, v in tr.full_state().items()}
        if exchange == "padded":
            tr.check_exchange()                                    # nothing was dropped at the default capacity
            assert tr.capacity(3 * B) < 3 * B                      # … and the buffers really are smaller than the worst case
            # a capacity that cannot hold the batch must be reported, not silently produce a wrong step
            from recommendit_b200 import RB200Error
            small = ShardedBPRTrainer(NU, NI, D, H, adam_mode=adam_mode, device="cpu", ops=OracleOps(), init=init, exchange="padded",
                                      capacity_factor=0.25)
            small.step(*[torch.from_numpy(a) for a in _batches(rank, 0)])
            try:
                small.check_exchange()
                raise AssertionError("overflow not reported")
            except RB200Error:
                pass
        if rank == 0:
            q.put((losses, full))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("adam_mode,exchange", [("dense", "exact"), ("rows", "exact"), ("rows", "padded"), ("dense", "padded")])
def test_sharded_step_equals_single_process(adam_mode, exchange):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, adam_mode, exchange, q)) for r in range(world)]
    for p in procs:
        p.start()
    losses, full = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process oracle on the concatenated global batch
    P = O.init_params(NU, NI, D, H, seed=3)
    S = O.AdamState()
    for step in range(2):
        parts = [_batches(r, step) for r in range(world)]
        gb = tuple(np.concatenate([pt[i] for pt in parts]) for i in range(5))
        if adam_mode == "dense":
            loss, _, _ = O.train_step(P, S, gb, lr=1e-2)
        else:   # touched-rows Adam: emulate by restoring untouched table rows and their moments
            before = {k: P[k].copy() for k in P}
            loss, _, _ = O.train_step(P, S, gb, lr=1e-2)
            for k, ids in (("user_tower.embedding.weight", gb[0]), ("item_tower.embedding.weight", np.concatenate([gb[1], gb[3]]))):
                un = np.setdiff1d(np.arange(P[k].shape[0]), np.unique(ids[ids != 0]))
                P[k][un] = before[k][un]; S.m[k][un] = 0; S.v[k][un] = 0
        assert abs(losses[step] - float(loss)) < 2e-6, (step, losses[step], float(loss))
    for k in O.PARAM_KEYS:
        assert np.abs(full[k] - P[k]).max() <= 0.25 * 1e-2, k


def test_route_plan_is_a_stable_partition():
    from recommendit_b200.sharded import make_route, shard_rows
    ids = torch.tensor([7, 0, 3, 8, 3, 12, 5, 0, 9])
    rt = make_route(ids, 4)
    owners = (ids % 4)[rt.perm]
    assert (owners[1:] >= owners[:-1]).all()                               # grouped by owner
    assert rt.send_counts == [int(((ids % 4) == r).sum()) for r in range(4)]
    assert torch.equal(ids[rt.perm] // 4, rt.local_rows)
    assert torch.equal(rt.perm[rt.inv], torch.arange(9))                   # inverse permutation
    for r in range(4):                                                     # stable inside a bucket
        idx = rt.perm[owners == r]
        assert (idx[1:] > idx[:-1]).all()
    assert [shard_rows(10, 4, r) for r in range(4)] == [3, 3, 2, 2]
    assert sum(shard_rows(10_000_001, 8, r) for r in range(8)) == 10_000_001


def _search_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from recommendit_b200.sharded import sharded_flat_search
        from tests.oracle_ops import flat_search_cpu, topk_merge_cpu
        rng = np.random.default_rng(0)
        x = V.normalize_rows(rng.standard_normal((3000, 32)).astype(np.float32))
        x[1500] = x[10]                                                    # an exact tie across shards
        qv = V.normalize_rows(rng.standard_normal((6, 32)).astype(np.float32)); qv[0] = x[10]
        lo, hi = (0, 1500) if rank == 0 else (1500, 3000)
        s, i = sharded_flat_search(torch.from_numpy(qv), torch.from_numpy(x[lo:hi]), 50, lo, search=flat_search_cpu,
                                   merge=topk_merge_cpu)
        if rank == 1:
            q.put((s.numpy(), i.numpy(), x, qv))
    finally:
        dist.destroy_process_group()


def test_sharded_flat_search_equals_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_search_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    s, i, x, qv = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    s_ref, i_ref = V.flat_search(qv, x, 50)
    np.testing.assert_array_equal(i, i_ref)            # incl. the cross-shard tie: lower id first
    np.testing.assert_allclose(s, s_ref, rtol=0, atol=1e-6)


class _OracleIVF:
    """CPU stand-in for FAISSIndex (the surface ShardedIVFIndex uses), arithmetic from oracle/ivf_oracle.py."""

    class _State:
        pass

    def __init__(self, d, nlist, nprobe):
        self.d, self.nlist, self.nprobe = d, nlist, nprobe
        self.index = None

    def build_ivf_index(self, x, ids, centroids=None):
        xn = V.normalize_rows(np.asarray(x, np.float32))
        cen = V.spherical_kmeans(xn, self.nlist) if centroids is None else np.asarray(centroids, np.float32)
        off, order = V.build_lists(V.assign(xn, cen), self.nlist)
        st = self._State()
        st.centroids, st.ntotal = torch.from_numpy(cen), len(ids)
        ids = np.asarray(ids, np.int64)

        def search_device(q, k, id_table=None):
            s, r = V.ivf_search(q.numpy(), cen, off, order, xn, self.nprobe, k)
            return torch.from_numpy(s), torch.from_numpy(np.where(r >= 0, ids[np.maximum(r, 0)], -1))
        st.search_device = search_device
        st.check_last_search = lambda: None          # (the device index reads its tensor-core error flag here)
        self.index, self._list_item_ids = st, None

    def set_n_probe(self, n):
        self.nprobe = n


def _ivf_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from recommendit_b200.sharded import ShardedIVFIndex
        from tests.oracle_ops import topk_merge_cpu
        rng = np.random.default_rng(5)
        x = V.normalize_rows(rng.standard_normal((2400, 32)).astype(np.float32))
        x[1900] = x[17]                                                    # an exact tie across the two shards
        ids = np.arange(1000, 3400)
        cen = V.spherical_kmeans(x, 12)
        qv = V.normalize_rows(rng.standard_normal((7, 32)).astype(np.float32)); qv[0] = x[17]
        lo, hi = (0, 1200) if rank == 0 else (1200, 2400)
        sh = ShardedIVFIndex(32, 12, 4, index_factory=lambda: _OracleIVF(32, 12, 4), merge=topk_merge_cpu)
        sh.build(x[lo:hi], ids[lo:hi], centroids=cen)
        assert sh.ntotal == 2400
        s, i = sh.search_device(torch.from_numpy(qv), 60)
        if rank == 1:
            q.put((s.numpy(), i.numpy(), x, ids, cen, qv))
    finally:
        dist.destroy_process_group()


def test_sharded_ivf_search_equals_unsharded():
    """Row-sharded IVFFlat (every rank: all centroids + its slice of every list; per-shard top-k → all-gather → merge) equals the
    unsharded IVF search on the same centroids, cross-shard tie included."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ivf_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    s, i, x, ids, cen, qv = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    off, order = V.build_lists(V.assign(x, cen), 12)
    s_ref, r_ref = V.ivf_search(qv, cen, off, order, x, 4, 60)
    np.testing.assert_array_equal(i, np.where(r_ref >= 0, ids[np.maximum(r_ref, 0)], -1))
    np.testing.assert_allclose(s, s_ref, rtol=0, atol=1e-6)
