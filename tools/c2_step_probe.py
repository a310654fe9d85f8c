"""C2 fused step probe (tools): eager steps for ncu, graph replays for timing.  usage: c2_step_probe.py [eager|graph] [steps]"""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import recommendit_b200 as R
from bench import N_USERS, N_ITEMS, D, H, DROPOUT, synth_batches
mode = sys.argv[1] if len(sys.argv) > 1 else "graph"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = R.TwoTowerModel(N_USERS, N_ITEMS, D, H, dropout=DROPOUT).to(dev).train()
tr = R.FusedBPRTrainer(model, use_cuda_graph=(mode == "graph"))
batches, _ = synth_batches(4)
res = [tr.pack_host(*b).clone().to(dev) for b in batches]
for i in range(5):
    tr.load_packed(res[i % 4]); tr.step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(steps):
    tr.load_packed(res[i % 4]); tr.step()
b.record(); torch.cuda.synchronize()
print("%s C2 step: %.4f ms, loss %.5f" % (mode, a.elapsed_time(b) / steps, float(tr.loss_dev)))
