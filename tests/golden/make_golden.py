"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports ``/root/reference/src/models/two_tower.py`` (PyTorch CPU, fp32), replays the reference
training-step body (``src/training/train_embeddings.py:183-192``) under fixed seeds and writes the
inputs and every observable output to ``tests/golden/*.npz``.  The fixtures travel to the GPU box;
``/root/reference`` does not, so nothing at test time reads it.

The IVF side of the reference (``faiss_index.py``) cannot be run here: ``faiss`` is not installed
(SURVEY.md F10), so there is no golden file for it — see ``oracle/ivf_oracle.py`` ("parity unpinned").
"""
import sys
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
sys.path.insert(0, str(REF))

from src.models.two_tower import TwoTowerModel  # noqa: E402  (the reference itself)


def sd_np(model):
    return {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}


def run_steps(name, n_users, n_items, D, H, B, steps, seed, dropout=0.0, with_pad=False, lr=1e-3):
    torch.manual_seed(seed)
    model = TwoTowerModel(n_users, n_items, embed_dim=D, hidden_dim=H, dropout=dropout)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=1e-5)
    out = {"meta": np.array([n_users, n_items, D, H, B, steps], dtype=np.int64),
           "dropout": np.array(dropout, dtype=np.float64), "lr": np.array(lr, dtype=np.float64)}
    for k, v in sd_np(model).items():
        out["init/" + k] = v
    # capture dropout masks (kept units) through forward hooks on the two nn.Dropout modules
    masks = []

    def hook(_m, inp, outp):
        x = inp[0]
        masks.append(torch.where(x > 0, (outp != 0), torch.ones_like(x, dtype=torch.bool)).numpy().copy())

    hs = [model.user_tower.mlp[2].register_forward_hook(hook),
          model.item_tower.mlp[2].register_forward_hook(hook)]
    g = torch.Generator().manual_seed(seed + 1)
    for s in range(steps):
        lo = 0 if with_pad else 1
        u = torch.randint(lo, n_users + 1, (B,), generator=g)
        p = torch.randint(lo, n_items + 1, (B,), generator=g)
        n = torch.randint(lo, n_items + 1, (B,), generator=g)
        pg = (torch.rand(B, 18, generator=g) < 0.15).float()
        ng = (torch.rand(B, 18, generator=g) < 0.15).float()
        masks.clear()
        # ---- the reference step body, train_embeddings.py:183-192 ----
        ue = model.user_tower(u)
        pe = model.item_tower(p, pg)
        ne = model.item_tower(n, ng)
        loss = model.bpr_loss(ue, pe, ne)
        opt.zero_grad()
        loss.backward()
        grads = {k: v.grad.detach().numpy().copy() for k, v in model.named_parameters()}
        total = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
        opt.step()
        # ---------------------------------------------------------------
        pre = f"step{s}/"
        out[pre + "user_ids"], out[pre + "pos_ids"], out[pre + "neg_ids"] = u.numpy(), p.numpy(), n.numpy()
        out[pre + "pos_genres"], out[pre + "neg_genres"] = pg.numpy(), ng.numpy()
        out[pre + "user_emb"], out[pre + "pos_emb"], out[pre + "neg_emb"] = (
            ue.detach().numpy().copy(), pe.detach().numpy().copy(), ne.detach().numpy().copy())
        out[pre + "loss"] = np.array(loss.item(), dtype=np.float32)
        out[pre + "total_norm"] = np.array(float(total), dtype=np.float32)
        if dropout > 0:
            out[pre + "mask_u"], out[pre + "mask_p"], out[pre + "mask_n"] = masks[0], masks[1], masks[2]
        for k, v in grads.items():
            out[pre + "grad/" + k] = v
        for k, v in sd_np(model).items():
            out[pre + "after/" + k] = v
    for h in hs:
        h.remove()
    np.savez_compressed(OUT / f"{name}.npz", **out)
    print(name, "loss", [float(out[f"step{s}/loss"]) for s in range(steps)])


def run_losses():
    torch.manual_seed(7)
    model = TwoTowerModel(10, 10, embed_dim=8, hidden_dim=8)
    out = {}
    for tag, B, D in (("a", 48, 32), ("b", 7, 64)):
        U = torch.nn.functional.normalize(torch.randn(B, D), dim=-1).requires_grad_(True)
        I = torch.nn.functional.normalize(torch.randn(B, D), dim=-1).requires_grad_(True)
        N = torch.nn.functional.normalize(torch.randn(B, D), dim=-1).requires_grad_(True)
        l = model.in_batch_bpr_loss(U, I)          # the literal Python loop, two_tower.py:143-160
        l.backward()
        out[f"{tag}/U"], out[f"{tag}/I"], out[f"{tag}/N"] = U.detach().numpy(), I.detach().numpy(), N.detach().numpy()
        out[f"{tag}/inbatch_loss"] = np.array(l.item(), np.float32)
        out[f"{tag}/inbatch_dU"], out[f"{tag}/inbatch_dI"] = U.grad.numpy().copy(), I.grad.numpy().copy()
        U.grad = None; I.grad = None
        l2 = model.bpr_loss(U, I, N)               # two_tower.py:127-129
        l2.backward()
        out[f"{tag}/bpr_loss"] = np.array(l2.item(), np.float32)
        out[f"{tag}/bpr_dU"], out[f"{tag}/bpr_dP"], out[f"{tag}/bpr_dN"] = (
            U.grad.numpy().copy(), I.grad.numpy().copy(), N.grad.numpy().copy())
    np.savez_compressed(OUT / "losses.npz", **out)
    print("losses", float(out["a/inbatch_loss"]), float(out["a/bpr_loss"]))


def run_inference():
    """get_user_embedding / get_item_embeddings (two_tower.py:166-196), reference test sizes."""
    torch.manual_seed(3)
    model = TwoTowerModel(100, 200, embed_dim=32, hidden_dim=64)
    out = {"init/" + k: v for k, v in sd_np(model).items()}
    rng = np.random.default_rng(5)
    genres = (rng.random((20, 18)) < 0.2).astype(np.float32)
    out["item_ids"] = np.arange(1, 21, dtype=np.int64)
    out["genres"] = genres
    out["item_embs"] = model.get_item_embeddings(list(range(1, 21)), genres, batch_size=8)
    out["user_emb_1"] = model.get_user_embedding(1)
    out["user_emb_100"] = model.get_user_embedding(100)
    np.savez_compressed(OUT / "inference.npz", **out)
    print("inference ok")


if __name__ == "__main__":
    # reference test-fixture sizes (tests/test_models.py:29-42)
    run_steps("tt_small", 100, 200, 32, 64, 16, steps=3, seed=11)
    # production widths, many duplicate ids + padding id 0 present
    run_steps("tt_dup", 300, 150, 64, 128, 256, steps=2, seed=12, with_pad=True)
    # dropout active, masks captured from the reference's own nn.Dropout
    run_steps("tt_dropout", 120, 90, 32, 64, 64, steps=2, seed=13, dropout=0.1)
    # D=128 (config C4 width), lr larger so the trajectory moves
    run_steps("tt_d128", 64, 80, 128, 128, 32, steps=2, seed=14, lr=1e-2)
    # production widths with dropout (exercises the tensor-core tower kernels with injected masks), ragged batch
    run_steps("tt_drop64", 700, 450, 64, 128, 200, steps=2, seed=15, dropout=0.1, with_pad=True)
    run_losses()
    run_inference()
