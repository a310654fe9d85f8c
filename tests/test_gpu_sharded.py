"""GPU tests of the sharded trainer with the real kernels (CudaOps): world 1 in-process, world 2 when two GPUs exist."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import two_tower_oracle as O
from tests.parity import batch_from_golden, dev, params_from_golden

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("adam_mode", ["dense", "rows"])
def test_sharded_world1_matches_golden(golden, adam_mode):
    from recommendit_b200.sharded import ShardedBPRTrainer
    g = golden("tt_d128")
    nu, ni, D, H = (int(v) for v in g["meta"][:4])
    init = {k: torch.from_numpy(g["init/" + k]) for k in O.PARAM_KEYS}
    tr = ShardedBPRTrainer(nu, ni, D, H, adam_mode=adam_mode, device="cuda", init=init, lr=float(g["lr"]))
    for s in range(2):
        b = batch_from_golden(g, s)
        loss = float(tr.step(*[dev(a) for a in b]))
        if adam_mode == "dense":
            assert abs(loss - float(g[f"step{s}/loss"])) < 5e-5
    if adam_mode == "dense":
        full = tr.full_state()
        for k in O.PARAM_KEYS:
            assert np.abs(full[k].cpu().numpy() - g["step1/after/" + k]).max() <= 0.5 * float(g["lr"]), k


WORKER = r'''
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, "%s")
from oracle import two_tower_oracle as O
from recommendit_b200.sharded import ShardedBPRTrainer, sharded_flat_search
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
NU, NI, D, H, B = 5003, 2999, 64, 128, 512
P = O.init_params(NU, NI, D, H, seed=5)
tr = ShardedBPRTrainer(NU, NI, D, H, adam_mode="dense", device=dev, init={k: torch.from_numpy(v) for k, v in P.items()}, lr=1e-2)
def batch(r, s):
    rng = np.random.default_rng(10 * s + r)
    return (rng.integers(0, NU + 1, B), rng.integers(0, NI + 1, B), (rng.random((B, 18)) < .2).astype(np.float32),
            rng.integers(0, NI + 1, B), (rng.random((B, 18)) < .2).astype(np.float32))
losses = [float(tr.step(*[torch.from_numpy(a).to(dev) for a in batch(rank, s)])) for s in range(2)]
full = tr.full_state()
# the same steps through the fixed-capacity exchange, eagerly and as ONE CUDA graph per step (NCCL collectives captured):
# 4 steps = 2 eager + capture + replay; must equal the exact exchange step for step (same kernels, same order of sums)
tr_e = ShardedBPRTrainer(NU, NI, D, H, adam_mode="dense", device=dev, init={k: torch.from_numpy(v) for k, v in P.items()}, lr=1e-2)
tr_g = ShardedBPRTrainer(NU, NI, D, H, adam_mode="dense", device=dev, init={k: torch.from_numpy(v) for k, v in P.items()}, lr=1e-2,
                         exchange="padded", capacity_factor=1.5, use_cuda_graph=True)
pad_diff = 0.0
for s in range(4):
    b = [torch.from_numpy(a).to(dev) for a in batch(rank, s)]
    le, lg = float(tr_e.step(*b)), float(tr_g.step(*b))
    pad_diff = max(pad_diff, abs(le - lg))
tr_g.check_exchange()
assert tr_g._graph is not None
fe, fg = tr_e.full_state(), tr_g.full_state()
pad_param_diff = max(float((fe[k] - fg[k]).abs().max()) for k in O.PARAM_KEYS)
tr_g.close()          # a live graph with NCCL nodes would hang destroy_process_group()
# … and through the peer-memory exchange (rows read from / gradients written to the owners' shards over NVLink, two cross-GPU
# barriers per step, symmetric memory): eager steps, capture, replays
tr_e2 = ShardedBPRTrainer(NU, NI, D, H, adam_mode="dense", device=dev, init={k: torch.from_numpy(v) for k, v in P.items()}, lr=1e-2)
tr_p = ShardedBPRTrainer(NU, NI, D, H, adam_mode="dense", device=dev, init={k: torch.from_numpy(v) for k, v in P.items()}, lr=1e-2,
                         exchange="p2p", capacity_factor=1.5, use_cuda_graph=True)
p2p_diff = 0.0
for s in range(5):
    b = [torch.from_numpy(a).to(dev) for a in batch(rank, s)]
    le, lp = float(tr_e2.step(*b)), float(tr_p.step(*b))
    p2p_diff = max(p2p_diff, abs(le - lp))
tr_p.check_exchange(); tr_p.check_ids()
assert tr_p._graph is not None
fe2, fp = tr_e2.full_state(), tr_p.full_state()
p2p_param_diff = max(float((fe2[k] - fp[k]).abs().max()) for k in O.PARAM_KEYS)
tr_p.close()
# sharded retrieval
g = torch.Generator().manual_seed(0)
x = torch.nn.functional.normalize(torch.randn(20000, 64, generator=g), dim=-1)
q = torch.nn.functional.normalize(torch.randn(9, 64, generator=g), dim=-1)
n_loc = 20000 // world
s, i = sharded_flat_search(q.to(dev), x[rank * n_loc:(rank + 1) * n_loc].contiguous().to(dev), 100, rank * n_loc)
# data-parallel replicas (small tables): must equal the single-process step on the concatenated batch
import recommendit_b200 as R
P2 = O.init_params(400, 300, 64, 128, seed=9)
m = R.TwoTowerModel(400, 300, 64, 128, dropout=0.0)
m.load_state_dict({k: torch.from_numpy(v) for k, v in P2.items()}); m.to(dev).train()
dp = R.DataParallelBPRTrainer(m, lr=1e-2, use_cuda_graph=True)
def batch2(r, s):
    rng = np.random.default_rng(77 * s + r)
    return (rng.integers(0, 401, 256), rng.integers(0, 301, 256), (rng.random((256, 18)) < .2).astype(np.float32),
            rng.integers(0, 301, 256), (rng.random((256, 18)) < .2).astype(np.float32))
dp_losses = [float(dp.step_host(*batch2(rank, s))) for s in range(4)]      # 2 eager + capture (incl. the all-reduce) + replay
dp_state = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
assert dp._graph is not None
dp.close()            # the captured step holds an NCCL node: release it before destroy_process_group()
# the same replicas with the all-reduce done over peer memory (two-shot kernel between two cross-GPU barriers, no NCCL in the step)
m2 = R.TwoTowerModel(400, 300, 64, 128, dropout=0.0)
m2.load_state_dict({k: torch.from_numpy(v) for k, v in O.init_params(400, 300, 64, 128, seed=9).items()}); m2.to(dev).train()
dp2 = R.DataParallelBPRTrainer(m2, lr=1e-2, use_cuda_graph=True, allreduce="p2p")
dp2_losses = [float(dp2.step_host(*batch2(rank, s))) for s in range(4)]
assert dp2._graph is not None and dp2._dp_hdl is not None
dp2_diff = max(abs(a - b) for a, b in zip(dp_losses, dp2_losses))
dp2_param_diff = max(float((v.detach().cpu() - torch.from_numpy(dp_state[k])).abs().max()) for k, v in m2.state_dict().items())
dp2.close()
# … and with the in-switch reduction switched off (two-shot kernel; where the fabric has no multicast object dp2 already used it)
import os
os.environ["RB200_DP_MULTIMEM"] = "0"
m3 = R.TwoTowerModel(400, 300, 64, 128, dropout=0.0)
m3.load_state_dict({k: torch.from_numpy(v) for k, v in O.init_params(400, 300, 64, 128, seed=9).items()}); m3.to(dev).train()
dp3 = R.DataParallelBPRTrainer(m3, lr=1e-2, use_cuda_graph=True, allreduce="p2p")
dp3_losses = [float(dp3.step_host(*batch2(rank, s))) for s in range(4)]
assert dp3._dp_mc == 0
dp3_diff = max(abs(a - b) for a, b in zip(dp_losses, dp3_losses))
dp3_param_diff = max(float((v.detach().cpu() - torch.from_numpy(dp_state[k])).abs().max()) for k, v in m3.state_dict().items())
dp_multimem_used = bool(dp2._dp_mc)
dp3.close()
del os.environ["RB200_DP_MULTIMEM"]
if rank == 0:
    S2 = O.AdamState(); dp_ref = []
    for st in range(4):
        parts = [batch2(r, st) for r in range(world)]
        gb = tuple(np.concatenate([p[k] for p in parts]) for k in range(5))
        dp_ref.append(float(O.train_step(P2, S2, gb, lr=1e-2)[0]))
    dp_err = max(float(np.abs(dp_state[k] - P2[k]).max()) for k in O.PARAM_KEYS)
if rank == 0:
    S = O.AdamState(); ref = []
    for st in range(2):
        parts = [batch(r, st) for r in range(world)]
        gb = tuple(np.concatenate([p[k] for p in parts]) for k in range(5))
        ref.append(float(O.train_step(P, S, gb, lr=1e-2)[0]))
    err = max(float(np.abs(full[k].cpu().numpy() - P[k]).max()) for k in O.PARAM_KEYS)
    from oracle import ivf_oracle as V
    s_ref, i_ref = V.flat_search(q.numpy(), x.numpy(), 100)
    V.assert_topk_equivalent(s.cpu().numpy(), i.cpu().numpy(), s_ref, i_ref)
    print("RESULT " + json.dumps({"losses": losses, "ref": ref, "max_param_err": err, "dp_losses": dp_losses, "dp_ref": dp_ref,
                                  "dp_err": dp_err, "pad_diff": pad_diff, "pad_param_diff": pad_param_diff,
                                  "p2p_diff": p2p_diff, "p2p_param_diff": p2p_param_diff,
                                  "dp2_diff": dp2_diff, "dp2_param_diff": dp2_param_diff, "dp3_diff": dp3_diff,
                                  "dp3_param_diff": dp3_param_diff, "dp_multimem_used": dp_multimem_used}))
dist.barrier(); dist.destroy_process_group()
'''


@pytest.mark.parametrize("exchange", ["padded", "p2p"])
def test_sharded_world1_padded_graph_equals_exact(golden, exchange):
    """exchange='padded' / 'p2p' + use_cuda_graph (the whole step one graph replay) against the exact exchange, world 1 (p2p at
    world 1 runs the same gather / push kernels on local pointers, without barriers)."""
    from recommendit_b200.sharded import ShardedBPRTrainer
    g = golden("tt_d128")
    nu, ni, D, H = (int(v) for v in g["meta"][:4])
    init = {k: torch.from_numpy(g["init/" + k]) for k in O.PARAM_KEYS}
    a = ShardedBPRTrainer(nu, ni, D, H, adam_mode="rows", device="cuda", init=init, lr=float(g["lr"]))
    b = ShardedBPRTrainer(nu, ni, D, H, adam_mode="rows", device="cuda", init=init, lr=float(g["lr"]), exchange=exchange,
                          use_cuda_graph=True)
    for s in range(5):
        batch = [dev(x) for x in batch_from_golden(g, s % 2)]
        la, lb = float(a.step(*batch)), float(b.step(*batch))
        assert abs(la - lb) <= 1e-6, (s, la, lb)
    assert b._graph is not None
    b.check_exchange()
    fa, fb = a.full_state(), b.full_state()
    for k in O.PARAM_KEYS:
        assert float((fa[k] - fb[k]).abs().max()) <= 1e-6, k


def test_sharded_train_epoch_from_pinned_batches_equals_step_by_step(golden):
    """ShardedBPRTrainer.train_epoch — batch i+1 uploaded on a copy stream under step i, losses read back asynchronously, one host
    synchronisation — must walk the same trajectory as calling step() on device batches one by one."""
    from recommendit_b200.sharded import ShardedBPRTrainer
    g = golden("tt_d128")
    nu, ni, D, H = (int(v) for v in g["meta"][:4])
    init = {k: torch.from_numpy(g["init/" + k]) for k in O.PARAM_KEYS}
    mk = lambda: ShardedBPRTrainer(nu, ni, D, H, adam_mode="rows", device="cuda", init=init, lr=float(g["lr"]), exchange="p2p",
                                   use_cuda_graph=True)
    a, b = mk(), mk()
    host = [[torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in batch_from_golden(g, s % 2)] for s in range(7)]
    step_losses = [float(a.step(*[t.cuda() for t in hb])) for hb in host]
    mean = b.train_epoch(host)
    assert abs(mean - float(np.mean(np.asarray(step_losses, np.float32)))) <= 1e-6, (mean, step_losses)
    fa, fb = a.full_state(), b.full_state()
    for k in O.PARAM_KEYS:
        assert torch.equal(fa[k], fb[k]), k


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_world2_nccl_matches_single_process_oracle(tmp_path):
    import json
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][0]
    r = json.loads(line[7:])
    assert np.allclose(r["losses"], r["ref"], atol=2e-5), r
    assert r["max_param_err"] <= 0.5 * 1e-2, r
    assert np.allclose(r["dp_losses"], r["dp_ref"], atol=2e-5), r
    assert r["dp_err"] <= 0.5 * 1e-2, r
    assert r["pad_diff"] <= 1e-6 and r["pad_param_diff"] <= 1e-6, r
    assert r["p2p_diff"] <= 1e-6 and r["p2p_param_diff"] <= 1e-6, r
    # peer-memory all-reduce of the data-parallel replicas vs the NCCL one (different summation orders of two addends: equal here)
    assert r["dp2_diff"] <= 1e-6 and r["dp2_param_diff"] <= 2e-5, r
    assert r["dp3_diff"] <= 1e-6 and r["dp3_param_diff"] <= 2e-5, r          # (dp2: multimem where available, dp3: two-shot)


@pytest.mark.parametrize("world,n_u,n_i", [(1, 100, 200), (2, 8192, 16384), (8, 5000, 10001), (3, 0, 77), (64, 300, 0)])
def test_route_plan_kernel_matches_tensor_reference(world, n_u, n_i):
    """rb200_route_plan (stable partition of the requests by owner) against the generic-tensor restatement."""
    from recommendit_b200.sharded import CudaOps, route_reference, shard_rows
    g = torch.Generator().manual_seed(world * 7 + n_u)
    u = torch.randint(0, 1_000_003, (n_u,), generator=g).cuda()
    i = torch.randint(0, 50_021, (n_i,), generator=g).cuda()
    nu_by_rank = torch.tensor([shard_rows(1_000_003, world, r) for r in range(world)], dtype=torch.int64, device="cuda")
    got = CudaOps().route(u, i, world, nu_by_rank)
    ref = route_reference(u, i, world, nu_by_rank)
    for a, b, name in zip(got, ref, ("perm", "inv", "local_rows", "send_counts")):
        assert torch.equal(a, b.to(a.dtype)), name


@pytest.mark.parametrize("world,n_u,n_i,cap", [(1, 100, 200, 300), (2, 8192, 16384, 13000), (8, 8192, 16384, 6160), (8, 5000, 10001, 1000),
                                               (3, 0, 77, 40), (64, 300, 0, 9)])
def test_route_plan_padded_kernel_matches_tensor_reference(world, n_u, n_i, cap):
    """rb200_route_plan_padded (fixed-capacity exchange plan) against the generic-tensor restatement, incl. buckets that overflow
    their capacity (the overflowing requests are counted; the slots of the others must still agree)."""
    from recommendit_b200.sharded import CudaOps, route_padded_reference, shard_rows
    g = torch.Generator().manual_seed(world * 11 + n_u)
    u = torch.randint(0, 1_000_003, (n_u,), generator=g).cuda()
    i = torch.randint(0, 50_021, (n_i,), generator=g).cuda()
    nu_by_rank = torch.tensor([shard_rows(1_000_003, world, r) for r in range(world)], dtype=torch.int64, device="cuda")
    of_k, of_r = torch.zeros(1, dtype=torch.int64, device="cuda"), torch.zeros(1, dtype=torch.int64, device="cuda")
    slot_k, rows_k = CudaOps().route_padded(u, i, world, nu_by_rank, cap, of_k)
    slot_r, rows_r = route_padded_reference(u, i, world, nu_by_rank, cap, of_r)
    assert int(of_k) == int(of_r)
    assert torch.equal(rows_k, rows_r)
    assert torch.equal(slot_k, slot_r)
    if world == 8 and cap == 1000:
        assert int(of_k) > 0                                   # this case really overflows


def test_padded_exchange_overflow_goes_to_the_dummy_slot():
    """A request that does not fit its bucket must cost that sample only (zero row in, gradient dropped), never touch another
    row, and must be reported — by check_exchange() and by the automatic check every `check_every` steps."""
    import recommendit_b200 as R
    from recommendit_b200.sharded import ShardedBPRTrainer
    NU, NI, D, H, B = 3000, 2000, 64, 128, 256
    P = O.init_params(NU, NI, D, H, seed=3)
    init = {k: torch.from_numpy(v) for k, v in P.items()}
    tr = ShardedBPRTrainer(NU, NI, D, H, adam_mode="rows", device="cuda", init=init, lr=1e-2, exchange="padded", capacity_factor=0.5,
                           check_every=2)
    C = tr.capacity(3 * B)
    assert C < 3 * B
    # distinct ids everywhere: requests in sample order [users | positives | negatives]; those at positions >= C overflow
    u = np.arange(1, B + 1); p = np.arange(1, B + 1); n = np.arange(B + 1, 2 * B + 1)
    z = np.zeros((B, 18), np.float32)
    before = tr.full_state()
    loss = float(tr.step(*[dev(a) for a in (u, p, z, n, z)]))
    assert np.isfinite(loss)
    after = tr.full_state()
    req_items = np.concatenate([p, n])
    fits = np.arange(B, 3 * B) < C                     # item requests that fit
    moved = (after["item_tower.embedding.weight"] - before["item_tower.embedding.weight"]).abs().sum(1).cpu().numpy() > 0
    assert moved[req_items[fits]].all()                # served requests update their rows
    assert not moved[req_items[~fits]].any()           # overflowed requests update nothing …
    untouched = np.setdiff1d(np.arange(NI + 1), req_items)
    assert not moved[untouched].any()                  # … and nobody else's row either
    with pytest.raises(R.RB200Error, match="exceeded the exchange capacity"):
        tr.check_exchange()
    tr.step(*[dev(a) for a in (u, p, z, n, z)])
    with pytest.raises(R.RB200Error, match="exceeded the exchange capacity"):
        tr.step(*[dev(a) for a in (u, p, z, n, z)])    # steps == 2: the automatic check fires


def test_sharded_dropout_is_seeded_and_graph_replays_draw_fresh_masks():
    from recommendit_b200.sharded import ShardedBPRTrainer
    NU, NI, D, H, B = 3000, 2000, 128, 128, 512
    P = O.init_params(NU, NI, D, H, seed=4)
    init = {k: torch.from_numpy(v) for k, v in P.items()}
    rng = np.random.default_rng(0)
    batch = [dev(a) for a in (rng.integers(1, NU + 1, B), rng.integers(1, NI + 1, B), (rng.random((B, 18)) < .2).astype(np.float32),
                              rng.integers(1, NI + 1, B), (rng.random((B, 18)) < .2).astype(np.float32))]

    def run(seed, graph, drop, lr=0.0):
        tr = ShardedBPRTrainer(NU, NI, D, H, adam_mode="rows", device="cuda", init=init, lr=lr, weight_decay=0.0, exchange="padded",
                               use_cuda_graph=graph, dropout=drop, seed=seed)
        return [float(tr.step(*batch)) for _ in range(5)]

    a, b, c, d = run(1, False, 0.3), run(1, True, 0.3), run(2, False, 0.3), run(1, False, 0.0)
    assert a == b                                        # replays read the step counter on the device: same masks as eager
    assert len(set(a)) == 5                              # lr = 0: only the masks change from step to step
    assert a != c and len(set(d)) == 1
