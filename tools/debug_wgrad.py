import sys; sys.path.insert(0, '.')
import numpy as np, torch
import recommendit_b200 as R
from oracle import two_tower_oracle as O
from tests.parity import rel_l2
torch.manual_seed(0)
nu, ni, B = 6040, 3952, 8192
rng = np.random.default_rng(2)
u, p, n = rng.integers(0, nu + 1, B), rng.integers(0, ni + 1, B), rng.integers(0, ni + 1, B)
table = (rng.random((ni + 1, 18)) < 0.1).astype(np.float32)
base = R.TwoTowerModel(nu, ni, 64, 128, dropout=0.0).cuda().train()
sd = {k: v.clone() for k, v in base.state_dict().items()}
P = {k: v.detach().cpu().numpy().astype(np.float64) for k, v in sd.items()}
l64, G64, _ = O.loss_and_grads(P, u, p, table[p].astype(np.float64), n, table[n].astype(np.float64))
H, D, Din = 128, 64, 82
for m in (0, 2, 1):
    model = R.TwoTowerModel(nu, ni, 64, 128, dropout=0.0).cuda().train(); model.load_state_dict(sd)
    tr = R.FusedBPRTrainer(model, use_cuda_graph=False, tower_mode=m, seed=5)
    tr.step_host(u, p, table[p], n, table[n])
    g = tr.views()["item_mlp_grad"].cpu().numpy()
    W1 = g[:H * Din].reshape(H, Din); b1 = g[H * Din:H * Din + H]; W2 = g[H * Din + H:H * Din + H + D * H].reshape(D, H)
    r1, r2 = G64["item_tower.mlp.0.weight"], G64["item_tower.mlp.3.weight"]
    print("mode", m, "W1 emb %.2e  W1 genre %.2e  b1 %.2e  W2 %.2e | norms W1emb %.3e genre %.3e W2 %.3e" % (
        rel_l2(W1[:, :64], r1[:, :64]), rel_l2(W1[:, 64:], r1[:, 64:]), rel_l2(b1, G64["item_tower.mlp.0.bias"]), rel_l2(W2, r2),
        np.linalg.norm(r1[:, :64]), np.linalg.norm(r1[:, 64:]), np.linalg.norm(r2)))
    # per genre column
    if m == 2:
        print("  per-genre-col err:", ["%.1e" % rel_l2(W1[:, 64 + k], r1[:, 64 + k]) for k in range(18)])
