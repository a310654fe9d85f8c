"""Pins oracle/two_tower_oracle.py against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU-only."""
import numpy as np
import pytest

from oracle import two_tower_oracle as O

CASES = ["tt_small", "tt_dup", "tt_dropout", "tt_d128", "tt_drop64"]


def _params(g, prefix, dtype=np.float32):
    return {k: g[prefix + k].astype(dtype) for k in O.PARAM_KEYS}


def _batch(g, s):
    p = f"step{s}/"
    return (g[p + "user_ids"], g[p + "pos_ids"], g[p + "pos_genres"], g[p + "neg_ids"], g[p + "neg_genres"])


def _masks(g, s):
    p = f"step{s}/"
    if p + "mask_u" not in g:
        return None
    return g[p + "mask_u"], g[p + "mask_p"], g[p + "mask_n"]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_trajectory_matches_reference(golden, case, dtype):
    g = golden(case)
    steps = int(g["meta"][5])
    drop_p = float(g["dropout"])
    lr = float(g["lr"])
    P = _params(g, "init/", dtype)
    S = O.AdamState()
    for s in range(steps):
        pre = f"step{s}/"
        # forward embeddings
        _, _, (u, p, n) = O.loss_and_grads(P, *_batch(g, s), masks=_masks(g, s), drop_p=drop_p)
        np.testing.assert_allclose(u, g[pre + "user_emb"], rtol=0, atol=2e-6)
        np.testing.assert_allclose(p, g[pre + "pos_emb"], rtol=0, atol=2e-6)
        np.testing.assert_allclose(n, g[pre + "neg_emb"], rtol=0, atol=2e-6)
        loss, G, _ = O.loss_and_grads(P, *_batch(g, s), masks=_masks(g, s), drop_p=drop_p)
        assert abs(float(loss) - float(g[pre + "loss"])) <= 1e-6 * max(1.0, abs(float(g[pre + "loss"])))
        for k in O.PARAM_KEYS:
            ref = g[pre + "grad/" + k]
            # parity metric: error relative to the tensor's L2 norm.  The fp64 oracle sits within
            # 4e-6 of the fp32 reference on every tensor (that is the reference's own rounding);
            # the fp32 oracle's sequential sums reach 1.7e-5 on the cancellation-heavy
            # item mlp.3.bias gradient (sum over 2B projected rows).
            # (condition number Σ|t|/|Σt| of that reduction is ≈270 on tt_dup, so any fp32
            # summation order lands 1e-5..5e-5 from the truth; the fp64 oracle does not.)
            tol = 1e-4 if k.endswith("mlp.3.bias") else 1e-5   # the reference's own fp32 sum is 1.4e-5 off the fp64 value there
            assert np.linalg.norm(G[k] - ref) <= tol * np.linalg.norm(ref) + 1e-12, (case, s, k)
        total = np.sqrt(sum((G[k].astype(np.float64) ** 2).sum() for k in O.PARAM_KEYS))
        assert abs(total - float(g[pre + "total_norm"])) <= 1e-5 * total
        # clip + Adam parity, isolated from gradient rounding: Adam divides by sqrt(v)+1e-8, so an
        # absolute gradient error of 1e-10 on a ~1e-9 gradient moves the update by ~10 % of lr.
        # Feed the reference's own gradients through the oracle's clip/Adam and continue the
        # trajectory from the reference's post-step parameters.
        Gref = {k: g[pre + "grad/" + k].astype(dtype) for k in O.PARAM_KEYS}
        coef, tot = O.clip_grad_norm([Gref[k] for k in O.PARAM_KEYS], 1.0)
        assert abs(float(tot) - float(g[pre + "total_norm"])) <= 2e-6 * float(tot)
        S.step += 1
        for k in O.PARAM_KEYS:
            if k not in S.m:
                S.m[k] = np.zeros_like(P[k]); S.v[k] = np.zeros_like(P[k])
            P[k], S.m[k], S.v[k] = O.adam_step(P[k], Gref[k] * coef, S.m[k], S.v[k], S.step, lr)
            ref = g[pre + "after/" + k]
            assert np.abs(P[k] - ref).max() <= 2e-3 * lr + 2e-7, (case, s, k, np.abs(P[k] - ref).max())
            P[k] = ref.astype(dtype)


@pytest.mark.parametrize("case", CASES)
def test_train_step_end_to_end_close(golden, case):
    """The chained oracle step (its own gradients) tracks the reference loss trajectory."""
    g = golden(case)
    steps = int(g["meta"][5]); drop_p = float(g["dropout"]); lr = float(g["lr"])
    P = _params(g, "init/"); S = O.AdamState()
    for s in range(steps):
        loss, _, _ = O.train_step(P, S, _batch(g, s), lr=lr, masks=_masks(g, s), drop_p=drop_p)
        assert abs(float(loss) - float(g[f"step{s}/loss"])) < 5e-5
    for k in O.PARAM_KEYS:
        moved = np.abs(g[f"step{steps-1}/after/" + k] - g["init/" + k]).max()
        assert np.abs(P[k] - g[f"step{steps-1}/after/" + k]).max() <= 0.5 * moved + 1e-7


def test_padding_row_gets_no_gradient(golden):
    g = golden("tt_dup")
    assert (g["step0/user_ids"] == 0).any() or (g["step0/pos_ids"] == 0).any()
    P = _params(g, "init/")
    _, G, _ = O.loss_and_grads(P, *_batch(g, 0))
    assert np.all(G["user_tower.embedding.weight"][0] == 0)
    assert np.all(G["item_tower.embedding.weight"][0] == 0)
    assert np.all(g["step0/grad/user_tower.embedding.weight"][0] == 0)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_losses_match_reference(golden, tag):
    g = golden("losses")
    U, I, N = g[f"{tag}/U"], g[f"{tag}/I"], g[f"{tag}/N"]
    l, dU, dI = O.in_batch_bpr_loss(U, I)
    assert abs(float(l) - float(g[f"{tag}/inbatch_loss"])) < 1e-6
    np.testing.assert_allclose(dU, g[f"{tag}/inbatch_dU"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(dI, g[f"{tag}/inbatch_dI"], rtol=1e-5, atol=1e-8)
    assert abs(float(O.in_batch_bpr_loss_loop(U, I)) - float(g[f"{tag}/inbatch_loss"])) < 1e-6
    l, dU, dP, dN = O.bpr_loss(U, I, N)
    assert abs(float(l) - float(g[f"{tag}/bpr_loss"])) < 1e-6
    np.testing.assert_allclose(dU, g[f"{tag}/bpr_dU"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(dP, g[f"{tag}/bpr_dP"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(dN, g[f"{tag}/bpr_dN"], rtol=1e-5, atol=1e-8)


def test_inference_helpers_match_reference(golden):
    g = golden("inference")
    P = _params(g, "init/")
    it = O._tower_args(P, "item")
    y, _ = O.tower_forward(it[0], g["item_ids"], g["genres"], *it[1:])
    np.testing.assert_allclose(y, g["item_embs"], atol=2e-6, rtol=0)
    ut = O._tower_args(P, "user")
    for uid in (1, 100):
        y, _ = O.tower_forward(ut[0], np.array([uid]), None, *ut[1:])
        np.testing.assert_allclose(y[0], g[f"user_emb_{uid}"], atol=2e-6, rtol=0)


def test_clamped_normalize_backward():
    """‖pre‖ < eps ⇒ output is pre/eps (pure scale) and the gradient is g/eps."""
    table = np.zeros((3, 4), np.float32)
    W1 = np.zeros((5, 4), np.float32); b1 = np.zeros(5, np.float32)
    W2 = np.zeros((4, 5), np.float32); b2 = np.zeros(4, np.float32)
    y, c = O.tower_forward(table, np.array([1, 2]), None, W1, b1, W2, b2)
    assert np.all(y == 0)
    dW1, db1, dW2, db2, dr = O.tower_backward(c, np.ones((2, 4), np.float32))
    assert np.allclose(db2, 2.0 / 1e-12)


@pytest.mark.parametrize("case", ["tt_small", "tt_d128"])
def test_torch_step_matches_reference_golden(golden, case):
    """oracle/torch_step.py (the CPU-baseline arm) reproduces the reference trajectory bit-for-bit-ish: it issues the
    same ATen calls."""
    import torch
    from oracle import torch_step as TS
    g = golden(case)
    T = TS.make_params({k: g["init/" + k] for k in O.PARAM_KEYS})
    opt = TS.make_optimizer(T, lr=float(g["lr"]))
    for s in range(int(g["meta"][5])):
        b = tuple(torch.from_numpy(a) for a in _batch(g, s))
        loss = TS.step(T, opt, b, dropout=0.0)
        assert abs(loss - float(g[f"step{s}/loss"])) < 1e-6
        for k in O.PARAM_KEYS:
            np.testing.assert_allclose(T[k].detach().numpy(), g[f"step{s}/after/" + k], rtol=0, atol=1e-6)
