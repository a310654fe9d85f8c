"""Probe (tools): padded exchange + CUDA graph on N ranks, progress printed step by step.  usage (torchrun): probe.py adam_mode D interleave"""
import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch, torch.distributed as dist
from oracle import two_tower_oracle as O
from recommendit_b200.sharded import ShardedBPRTrainer
adam_mode, D, interleave = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
NU, NI, H, B = 5003, 2999, 128, 512
P = O.init_params(NU, NI, D, H, seed=5)
init = {k: torch.from_numpy(v) for k, v in P.items()}
def say(*a):
    print(f"[r{rank} {time.time() % 1000:.2f}]", *a, flush=True)
def batch(r, s):
    rng = np.random.default_rng(10 * s + r)
    return (rng.integers(0, NU + 1, B), rng.integers(0, NI + 1, B), (rng.random((B, 18)) < .2).astype(np.float32),
            rng.integers(0, NI + 1, B), (rng.random((B, 18)) < .2).astype(np.float32))
tr_e = ShardedBPRTrainer(NU, NI, D, H, adam_mode=adam_mode, device=dev, init=init, lr=1e-2) if interleave else None
tr_g = ShardedBPRTrainer(NU, NI, D, H, adam_mode=adam_mode, device=dev, init=init, lr=1e-2, exchange="padded", capacity_factor=1.5,
                         use_cuda_graph=True)
for s in range(5):
    b = [torch.from_numpy(a).to(dev) for a in batch(rank, s)]
    if tr_e is not None:
        le = float(tr_e.step(*b)); say("exact step", s, le)
    lg = float(tr_g.step(*b)); say("padded step", s, lg, "graph" if tr_g._graph is not None else "eager")
tr_g.check_exchange()
say("full_state")
fg = tr_g.full_state()
say("done")
tr_g.close()
dist.barrier(); dist.destroy_process_group()
