"""Small invocations of the round-2 kernels for compute-sanitizer (tools): exhaustive search (stream / filter kernels), IVF search
(persistent list scan), sharded step at world 1 (plan / apply scatter, peer kernels on local pointers)."""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import recommendit_b200 as R
from oracle import ivf_oracle as V
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.nn.functional.normalize(torch.randn(40000, 64, device="cuda", generator=g), dim=-1)
for nq in (5, 200):
    q = torch.nn.functional.normalize(torch.randn(nq, 64, device="cuda", generator=g), dim=-1)
    s, i = R.flat_search(q, x, 50)
    ref = torch.topk(q @ x.T, 50, dim=1)
    print("flat", nq, float((i == ref.indices).float().mean()))
rng = np.random.default_rng(0)
xs = V.normalize_rows(rng.standard_normal((6000, 64)).astype(np.float32))
cen = V.spherical_kmeans(xs, 32)
idx = R.FAISSIndex(64, 32, 8)
idx.build_ivf_index(xs, list(range(1, 6001)), centroids=cen)
qs = V.normalize_rows(rng.standard_normal((300, 64)).astype(np.float32))
s, ids = idx.batch_search(qs, 40)
off, order = V.build_lists(V.assign(xs, cen), 32)
s_ref, i_ref = V.ivf_search(qs, cen, off, order, xs, 8, 40)
V.assert_topk_equivalent(s, ids, s_ref, np.where(i_ref >= 0, i_ref + 1, -1))
print("ivf ok")
from recommendit_b200.sharded import ShardedBPRTrainer
tr = ShardedBPRTrainer(5000, 800, 128, 128, 18, adam_mode="rows", device="cuda", seed=3, exchange="p2p", use_cuda_graph=False, dropout=0.1)
B = 2048
for st in range(2):
    u = torch.from_numpy((rng.zipf(1.05, B) - 1) % 5000 + 1).cuda()
    p, n = torch.from_numpy(rng.integers(1, 801, B)).cuda(), torch.from_numpy(rng.integers(1, 801, B)).cuda()
    pg = torch.from_numpy((rng.random((B, 18)) < 0.1).astype(np.float32)).cuda()
    ng = torch.from_numpy((rng.random((B, 18)) < 0.1).astype(np.float32)).cuda()
    print("sharded loss", float(tr.step(u, p, pg, n, ng)))
torch.cuda.synchronize()
print("SANITY-DONE")
