import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
import recommendit_b200 as R
from bench import synth_batches, N_USERS, N_ITEMS, D, H
dev = torch.device("cuda")
model = R.TwoTowerModel(N_USERS, N_ITEMS, D, H, dropout=0.1).to(dev).train()
batches, _ = synth_batches(8)
tr = R.FusedBPRTrainer(model)
pinned = [tr.pack_host(*b).clone().pin_memory() for b in batches]
for i in range(5):
    tr.load_packed(pinned[i % 8]); tr.step()
torch.cuda.synchronize()
def t(fn, n=50):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): fn(i)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("h2d+step+item   %.3f ms" % t(lambda i: (tr.load_packed(pinned[i % 8]), tr.step().item())))
print("step+item       %.3f ms" % t(lambda i: tr.step().item()))
print("step+sync       %.3f ms" % t(lambda i: (tr.step(), torch.cuda.synchronize())))
print("step (no sync)  %.3f ms" % t(lambda i: tr.step()))
print("replay+sync     %.3f ms" % t(lambda i: (tr._graph.replay(), torch.cuda.synchronize())))
print("h2d+sync        %.3f ms" % t(lambda i: (tr.load_packed(pinned[i % 8]), torch.cuda.synchronize())))
tr2 = R.FusedBPRTrainer(model, use_cuda_graph=False); tr2._alloc(8192)
tr2.load_packed(pinned[0]); tr2.step()
print("eager step+sync %.3f ms" % t(lambda i: (tr2.step(), torch.cuda.synchronize())))
flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
def timed(pre, n=50):
    tot = 0.0
    for i in range(n):
        pre(); torch.cuda.synchronize(); t0 = time.perf_counter()
        tr.load_packed(pinned[i % 8]); tr.step().item()
        tot += time.perf_counter() - t0
    return tot / n * 1e3
print("sync only       %.3f ms" % timed(lambda: None))
print("flush+sync      %.3f ms" % timed(lambda: flush.add_(1)))
small = torch.zeros(1 << 20, dtype=torch.uint8, device=dev)
print("small kernel+sync %.3f ms" % timed(lambda: small.add_(1)))
import time as _t
print("sleep 2ms + sync %.3f ms" % timed(lambda: _t.sleep(0.002)))
# event-timed step after flush (as in bench value loop)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
for a, b in ev:
    flush.add_(1); a.record(); tr.load_packed(pinned[0]); tr.step(); b.record()
torch.cuda.synchronize()
print("event-timed after flush %.3f ms" % (sum(a.elapsed_time(b) for a, b in ev) / 20))
