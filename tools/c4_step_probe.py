"""One rank of the C4 sharded step at world 1 (tools): eager (so that an ncu launch list shows every kernel) and as a CUDA-graph replay
(device time without launch overhead).  usage: c4_step_probe.py [graph|eager] [p2p|padded]"""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from recommendit_b200.sharded import ShardedBPRTrainer
mode = sys.argv[1] if len(sys.argv) > 1 else "graph"
dev = torch.device("cuda", 0)
NU4, NI4, D4, B = 10_000_000 // 8, 1_000_000 // 8, 128, 8192          # one rank's share of the C4 tables at world 8
exch = sys.argv[2] if len(sys.argv) > 2 else "p2p"
tr = ShardedBPRTrainer(NU4, NI4, D4, 128, 18, adam_mode="rows", device=dev, seed=11, exchange=exch, capacity_factor=16.0 if exch == "padded" else 2.0,
                       use_cuda_graph=(mode == "graph"), dropout=0.1)
rng = np.random.default_rng(0)
bs = []
for _ in range(4):
    u = (rng.zipf(1.05, B) - 1) % NU4 + 1
    p, n = rng.integers(1, NI4 + 1, B), rng.integers(1, NI4 + 1, B)
    pg, ng = (rng.random((B, 18)) < 0.092).astype(np.float32), (rng.random((B, 18)) < 0.092).astype(np.float32)
    bs.append(tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (u, p, pg, n, ng)))
for i in range(4):
    tr.step(*bs[i])
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(20):
    tr.step(*bs[i % 4])
b.record(); torch.cuda.synchronize()
print("%s step, world 1, dropout 0.1: %.3f ms, loss %.5f" % (mode, a.elapsed_time(b) / 20, float(tr.step(*bs[0]))), flush=True)
tr.check_exchange()
