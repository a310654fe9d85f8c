"""Timing of rb200_flat_search on one C5 shard (tools; not part of the product).  usage: flat_prof.py [rows] [nq]"""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import recommendit_b200 as R
import bench

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
nqs = [int(a) for a in sys.argv[2:]] or [4096, 64, 1]
dev = torch.device("cuda", 0)
x = bench.make_flat_shard(dev, rows, 13)
g = torch.Generator(device=dev).manual_seed(14)
for nq in nqs:
    q = torch.nn.functional.normalize(torch.randn(nq, 64, device=dev, generator=g), dim=-1)
    R.flat_search(q, x, 500); torch.cuda.synchronize()
    ms = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); s, i = R.flat_search(q, x, 500); b.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    t = float(np.median(ms))
    sub = min(rows, 2_000_000)
    s1, i1 = R.flat_search(q[:32], x[:sub], 500)
    ref = torch.topk(q[:32].double() @ x[:sub].double().T, 500, dim=1)
    print(f"rows {rows} nq {nq}: {t:.3f} ms  {nq / t * 1e3:.0f} q/s  {2.0 * nq * rows * 64 / t / 1e9:.1f} logical TFLOP/s  "
          f"ids==fp64 topk on {sub} rows: {float((i1 == ref.indices).float().mean()):.6f}  "
          f"max score err {float((s1.double() - ref.values).abs().max()):.2e}", flush=True)
