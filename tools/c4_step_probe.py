"""One rank of the C4 sharded step at world 1 (tools): eager padded step, so that an ncu launch list shows every kernel."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from recommendit_b200.sharded import ShardedBPRTrainer
dev = torch.device("cuda", 0)
NU4, NI4, D4, B = 10_000_000 // 8, 1_000_000 // 8, 128, 8192          # one rank's share of the C4 tables at world 8
tr = ShardedBPRTrainer(NU4, NI4, D4, 128, 18, adam_mode="rows", device=dev, seed=11, exchange="padded", capacity_factor=16.0)
rng = np.random.default_rng(0)
bs = []
for _ in range(4):
    u = (rng.zipf(1.05, B) - 1) % NU4 + 1
    p, n = rng.integers(1, NI4 + 1, B), rng.integers(1, NI4 + 1, B)
    pg, ng = (rng.random((B, 18)) < 0.092).astype(np.float32), (rng.random((B, 18)) < 0.092).astype(np.float32)
    bs.append(tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (u, p, pg, n, ng)))
for i in range(3):
    tr.step(*bs[i])
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(20):
    tr.step(*bs[i % 4])
b.record(); torch.cuda.synchronize()
print("eager padded step, world 1: %.3f ms" % (a.elapsed_time(b) / 20), flush=True)
