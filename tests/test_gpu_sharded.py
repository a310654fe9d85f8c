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
# sharded retrieval
g = torch.Generator().manual_seed(0)
x = torch.nn.functional.normalize(torch.randn(20000, 64, generator=g), dim=-1)
q = torch.nn.functional.normalize(torch.randn(9, 64, generator=g), dim=-1)
n_loc = 20000 // world
s, i = sharded_flat_search(q.to(dev), x[rank * n_loc:(rank + 1) * n_loc].contiguous().to(dev), 100, rank * n_loc)
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
    print("RESULT " + json.dumps({"losses": losses, "ref": ref, "max_param_err": err}))
dist.barrier(); dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_world2_nccl_matches_single_process_oracle(tmp_path):
    import json
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][0]
    r = json.loads(line[7:])
    assert np.allclose(r["losses"], r["ref"], atol=2e-5), r
    assert r["max_param_err"] <= 0.5 * 1e-2, r
