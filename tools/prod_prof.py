"""Device producer: in-graph sampling vs one sampling launch per step (tools; not part of the product)."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import recommendit_b200 as R
from bench import N_USERS, N_ITEMS, D, H, B, DROPOUT

dev = torch.device("cuda", 0)
rng = np.random.default_rng(20240601)
n = 1_000_209
catalog = np.sort(rng.choice(np.arange(1, N_ITEMS + 1), 3883, replace=False)).astype(np.int64)
u = rng.integers(1, N_USERS + 1, n); i = catalog[rng.integers(0, 3706, n)]
r = rng.choice([1, 2, 3, 4, 5], n, p=[0.056, 0.107, 0.261, 0.349, 0.227]).astype(np.float64)
genres = (rng.random((N_ITEMS + 1, 18)) < 0.092).astype(np.float32)
for in_graph in (True, False):
    prod = R.DeviceBatchProducer(u, i, r, catalog, N_USERS, seed=1)
    torch.manual_seed(0)
    model = R.TwoTowerModel(N_USERS, N_ITEMS, D, H, dropout=DROPOUT).to(dev).train()
    tr = R.FusedBPRTrainer(model, item_extra_table=torch.from_numpy(genres))
    nb = prod.batches_per_epoch(B)
    prod.train_epoch(tr, B, 0, in_graph=in_graph)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for e in range(1, 6):
        prod.train_epoch(tr, B, e, in_graph=in_graph)
    dt = time.perf_counter() - t0
    # pure replays, no per-step torch op
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(nb):
        tr.step()
    b.record(); torch.cuda.synchronize()
    print(f"in_graph={in_graph}: epoch loop {dt / (5 * nb) * 1e3:.4f} ms/step ({5 * nb * B / dt / 1e6:.1f} M samples/s); "
          f"bare replays {a.elapsed_time(b) / nb:.4f} ms/step", flush=True)
