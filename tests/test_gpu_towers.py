"""GPU parity: tower forward/backward, BPR losses and the dense embedding gradient, through the drop-in Python
classes (→ ctypes → C ABI → sm_100a kernels), against golden vectors from the reference and the fp64 oracle."""
import numpy as np
import pytest
import torch

from oracle import two_tower_oracle as O
from tests.parity import (batch_from_golden, dev, grad_tol, masks_from_golden, model_from_golden, params_from_golden,
                          rel_l2)

pytestmark = pytest.mark.gpu
CASES = ["tt_small", "tt_dup", "tt_dropout", "tt_d128", "tt_drop64"]


def _run_reference_step_body(model, g, s):
    """The reference's step body (train_embeddings.py:183-190) on the drop-in model, up to backward."""
    u, p, pg, n, ng = batch_from_golden(g, s)
    masks = masks_from_golden(g, s)
    mk = [None] * 3 if masks is None else [dev(m) for m in masks]
    user_emb = model.user_tower(dev(u), keep_mask=mk[0])
    pos_emb = model.item_tower(dev(p), dev(pg), keep_mask=mk[1])
    neg_emb = model.item_tower(dev(n), dev(ng), keep_mask=mk[2])
    loss = model.bpr_loss(user_emb, pos_emb, neg_emb)
    model.zero_grad()
    loss.backward()
    return user_emb, pos_emb, neg_emb, loss


@pytest.mark.parametrize("case", CASES)
def test_forward_backward_match_reference_golden(golden, case):
    g = golden(case)
    model = model_from_golden(g).train()
    ue, pe, ne, loss = _run_reference_step_body(model, g, 0)
    np.testing.assert_allclose(ue.detach().cpu().numpy(), g["step0/user_emb"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(pe.detach().cpu().numpy(), g["step0/pos_emb"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(ne.detach().cpu().numpy(), g["step0/neg_emb"], atol=2e-6, rtol=0)
    assert abs(loss.item() - float(g["step0/loss"])) <= 1e-6
    # gradients: against the fp64 oracle (tight) and against the reference's fp32 values
    P = params_from_golden(g)
    _, G64, _ = O.loss_and_grads(P, *batch_from_golden(g, 0), masks=masks_from_golden(g, 0), drop_p=float(g["dropout"]))
    for k, prm in model.named_parameters():
        got = prm.grad.detach().cpu().numpy()
        assert rel_l2(got, G64[k]) <= grad_tol(k), (case, k, rel_l2(got, G64[k]))
        assert rel_l2(got, g["step0/grad/" + k]) <= 3 * grad_tol(k), (case, k)


def test_padding_row_gets_no_gradient_and_duplicates_sum(golden):
    g = golden("tt_dup")
    model = model_from_golden(g).train()
    _run_reference_step_body(model, g, 0)
    gu = model.user_tower.embedding.weight.grad
    gi = model.item_tower.embedding.weight.grad
    assert (g["step0/user_ids"] == 0).any()
    assert torch.count_nonzero(gu[0]) == 0 and torch.count_nonzero(gi[0]) == 0
    touched = np.unique(np.concatenate([g["step0/pos_ids"], g["step0/neg_ids"]]))
    untouched = np.setdiff1d(np.arange(gi.shape[0]), touched)
    assert torch.count_nonzero(gi[dev(untouched)]) == 0
    assert gu.is_contiguous() and not gu.is_sparse      # dense grads: torch.optim.Adam rejects sparse ones


def test_backward_is_deterministic(golden):
    g = golden("tt_dup")
    model = model_from_golden(g).train()
    _run_reference_step_body(model, g, 0)
    first = {k: p.grad.clone() for k, p in model.named_parameters()}
    for _ in range(3):
        _run_reference_step_body(model, g, 0)
        for k, p in model.named_parameters():
            assert torch.equal(p.grad, first[k]), k


def test_reference_training_loop_with_torch_adam_tracks_golden(golden):
    """The unchanged caller owns torch.optim.Adam + clip_grad_norm_ (train_embeddings.py:160,189-192)."""
    for case in ("tt_small", "tt_d128"):
        g = golden(case)
        model = model_from_golden(g).train()
        opt = torch.optim.Adam(model.parameters(), lr=float(g["lr"]), weight_decay=1e-5)
        for s in range(int(g["meta"][5])):
            u, p, pg, n, ng = (dev(a) for a in batch_from_golden(g, s))
            loss = model.bpr_loss(model.user_tower(u), model.item_tower(p, pg), model.item_tower(n, ng))
            opt.zero_grad()
            loss.backward()
            total = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
            opt.step()
            assert abs(loss.item() - float(g[f"step{s}/loss"])) < 5e-5, (case, s)
            assert abs(float(total) - float(g[f"step{s}/total_norm"])) <= 2e-5 * float(total)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_losses_match_reference_golden(golden, tag):
    import recommendit_b200 as R
    g = golden("losses")
    model = R.TwoTowerModel(4, 4, 32, 64).cuda()
    U, I, N = (dev(g[f"{tag}/{x}"]).requires_grad_(True) for x in "UIN")
    loss = model.bpr_loss(U, I, N)
    loss.backward()
    assert abs(loss.item() - float(g[f"{tag}/bpr_loss"])) < 1e-6
    for t, name in ((U, "bpr_dU"), (I, "bpr_dP"), (N, "bpr_dN")):
        np.testing.assert_allclose(t.grad.cpu().numpy(), g[f"{tag}/{name}"], rtol=1e-5, atol=1e-8)
    U.grad = I.grad = None
    loss = model.in_batch_bpr_loss(U, I, mode=0)
    loss.backward()
    assert abs(loss.item() - float(g[f"{tag}/inbatch_loss"])) < 1e-6
    np.testing.assert_allclose(U.grad.cpu().numpy(), g[f"{tag}/inbatch_dU"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(I.grad.cpu().numpy(), g[f"{tag}/inbatch_dI"], rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("B,D", [(1, 32), (2, 64), (63, 64), (64, 64), (65, 128), (200, 64), (1000, 32)])
def test_inbatch_loss_ragged_sizes_vs_oracle(B, D):
    import recommendit_b200 as R
    rng = np.random.default_rng(B * 1000 + D)
    U = rng.standard_normal((B, D)); U /= np.linalg.norm(U, axis=1, keepdims=True)
    I = rng.standard_normal((B, D)); I /= np.linalg.norm(I, axis=1, keepdims=True)
    model = R.TwoTowerModel(4, 4, 32, 64).cuda()
    Ut, It = dev(U, torch.float32).requires_grad_(True), dev(I, torch.float32).requires_grad_(True)
    loss = model.in_batch_bpr_loss(Ut, It, mode=0)
    if B == 1:
        assert np.isnan(loss.item())          # mean over an empty set of negatives, as in the reference loop
        return
    loss.backward()
    l64, dU, dI = O.in_batch_bpr_loss(U, I)
    assert abs(loss.item() - float(l64)) <= 1e-6
    assert rel_l2(Ut.grad.cpu().numpy(), dU) <= 1e-5
    assert rel_l2(It.grad.cpu().numpy(), dI) <= 1e-5


@pytest.mark.parametrize("B", [1, 5, 64, 77, 1024])
@pytest.mark.parametrize("D,H", [(32, 64), (64, 128), (128, 128), (32, 256)])
def test_tower_shapes_vs_oracle(B, D, H):
    import recommendit_b200 as R
    torch.manual_seed(B + D + H)
    model = R.TwoTowerModel(50, 70, D, H, dropout=0.0).cuda().train()
    P = {k: v.detach().cpu().numpy().astype(np.float64) for k, v in model.state_dict().items()}
    rng = np.random.default_rng(B)
    u, p = rng.integers(0, 51, B), rng.integers(0, 71, B)
    pg = (rng.random((B, 18)) < 0.2).astype(np.float32)
    ue, pe = model(dev(u), dev(p), dev(pg))
    ut, it = O._tower_args(P, "user"), O._tower_args(P, "item")
    yu, cu = O.tower_forward(ut[0], u, None, *ut[1:])
    yp, cp = O.tower_forward(it[0], p, pg.astype(np.float64), *it[1:])
    np.testing.assert_allclose(ue.detach().cpu().numpy(), yu, atol=2e-6, rtol=0)
    np.testing.assert_allclose(pe.detach().cpu().numpy(), yp, atol=2e-6, rtol=0)
    w = dev(rng.standard_normal((B, D)), torch.float32)
    ((ue * w).sum() + (pe * w).sum()).backward()
    dW1, db1, dW2, db2, dr = O.tower_backward(cp, w.cpu().numpy().astype(np.float64))
    t = model.item_tower
    assert rel_l2(t.mlp[0].weight.grad.cpu().numpy(), dW1) <= 1e-5
    assert rel_l2(t.mlp[0].bias.grad.cpu().numpy(), db1) <= 1e-5
    assert rel_l2(t.mlp[3].weight.grad.cpu().numpy(), dW2) <= 1e-5
    assert rel_l2(t.mlp[3].bias.grad.cpu().numpy(), db2) <= 1e-4
    assert rel_l2(t.embedding.weight.grad.cpu().numpy(), O.embedding_dense_backward(p, dr, 71)) <= 1e-5


def test_in_kernel_dropout_statistics():
    """Philox mask: keep rate ≈ 1-p, inverted scaling, different masks per call, eval() is deterministic."""
    import recommendit_b200 as R
    torch.manual_seed(1)
    model = R.TwoTowerModel(1000, 10, 64, 128, dropout=0.1).cuda().train()
    ids = torch.arange(1, 1001, device="cuda")
    a = model.user_tower(ids)
    b = model.user_tower(ids)
    assert not torch.equal(a, b)
    assert torch.allclose(a.norm(dim=-1), torch.ones(1000, device="cuda"), atol=1e-5)
    model.eval()
    c, d = model.user_tower(ids), model.user_tower(ids)
    assert torch.equal(c, d)
    # keep-rate through the saved hidden activations
    from recommendit_b200.two_tower import _TowerFn
    model.train()
    t = model.user_tower
    w = t.embedding.weight.detach().clone().requires_grad_(True)
    out = _TowerFn.apply(ids, None, w, t.mlp[0].weight, t.mlp[0].bias, t.mlp[3].weight, t.mlp[3].bias, 0.5, 7, 0, None)
    hid = out.grad_fn.saved_tensors[6]
    ref = torch.relu(torch.nn.functional.linear(w[ids], t.mlp[0].weight, t.mlp[0].bias))
    live = ref > 1e-6
    kept = (hid > 0) & live
    rate = kept.sum().item() / live.sum().item()
    assert 0.48 < rate < 0.52, rate
    assert torch.allclose(hid[kept], 2.0 * ref[kept], rtol=1e-5, atol=1e-6)


def test_out_of_range_and_cpu_inputs_raise():
    import recommendit_b200 as R
    model = R.TwoTowerModel(10, 10, 32, 64).cuda()
    with pytest.raises(R.RB200Error):
        model.user_tower(torch.tensor([1, 2]))                 # CPU ids on a CUDA model
    with pytest.raises(R.RB200Error):
        R.TwoTowerModel(10, 10, 48, 64).cuda().user_tower(torch.tensor([1], device="cuda"))   # unsupported width


def test_width_outside_the_shared_memory_budget_is_refused():
    """D=64, H=256 item tower needs 242 KB of shared memory (> 227 KB): a clear error, not a wrong answer."""
    import recommendit_b200 as R
    m = R.TwoTowerModel(10, 10, 64, 256).cuda()
    with pytest.raises(R.RB200Error, match="shared memory"):
        m.item_tower(torch.tensor([1], device="cuda"), torch.zeros(1, 18, device="cuda"))
