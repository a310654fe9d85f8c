"""In-batch BPR on the tensor cores (csrc/inbatch_tc.cu) against the fp64 oracle and the reference's golden vectors.

Reference: src/models/two_tower.py:132-160 (the Python loop over users)."""
import numpy as np
import pytest
import torch

from oracle import two_tower_oracle as O
from tests.parity import dev, rel_l2

pytestmark = pytest.mark.gpu


def _unit(rng, B, D):
    x = rng.standard_normal((B, D))
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def _run(U, I, mode, grad=True):
    import recommendit_b200 as R
    model = R.TwoTowerModel(4, 4, 32, 64).cuda()
    Ut, It = dev(U, torch.float32).requires_grad_(grad), dev(I, torch.float32).requires_grad_(grad)
    loss = model.in_batch_bpr_loss(Ut, It, mode=mode)
    if grad:
        loss.backward()
        return loss.item(), Ut.grad.cpu().numpy(), It.grad.cpu().numpy()
    return loss.item(), None, None


@pytest.mark.parametrize("B", [2, 7, 63, 64, 65, 128, 129, 200, 1000, 2500])
def test_inbatch_3xtf32_vs_oracle(B):
    rng = np.random.default_rng(B)
    U, I = _unit(rng, B, 64), _unit(rng, B, 64)
    loss, dU, dI = _run(U, I, 2)
    l64, rU, rI = O.in_batch_bpr_loss(U, I)
    assert abs(loss - float(l64)) <= 1e-6
    assert rel_l2(dU, rU) <= 1e-5 and rel_l2(dI, rI) <= 1e-5
    # the SIMT kernel and the tensor-core kernel agree with each other as well
    loss0, dU0, dI0 = _run(U, I, 0)
    assert abs(loss - loss0) <= 1e-6 and rel_l2(dU, dU0) <= 1e-5 and rel_l2(dI, dI0) <= 1e-5


def test_inbatch_3xtf32_reference_golden(golden):
    g = golden("losses")                     # case b: B=7, D=64, produced by the reference's own loop
    loss, dU, dI = _run(g["b/U"], g["b/I"], 2)
    assert abs(loss - float(g["b/inbatch_loss"])) < 1e-6
    np.testing.assert_allclose(dU, g["b/inbatch_dU"], rtol=1e-4, atol=2e-8)
    np.testing.assert_allclose(dI, g["b/inbatch_dI"], rtol=1e-4, atol=2e-8)


def test_inbatch_correlated_embeddings():
    """Trained-model regime: positives score high (S_ii ≈ 0.9), many near-duplicate items."""
    rng = np.random.default_rng(3)
    B = 777
    U = _unit(rng, B, 64)
    I = U + 0.3 * rng.standard_normal((B, 64)); I /= np.linalg.norm(I, axis=1, keepdims=True)
    I[10] = I[11]
    loss, dU, dI = _run(U, I, 2)
    l64, rU, rI = O.in_batch_bpr_loss(U, I)
    assert abs(loss - float(l64)) <= 1e-6
    assert rel_l2(dU, rU) <= 1e-5 and rel_l2(dI, rI) <= 1e-5


def test_inbatch_untrained_regime_all_embeddings_alike():
    """Untrained towers: every embedding is close to one common direction, so dU = Σ_j G_ij I_j − r_i I_i is a small
    difference of large terms — the kernel centres the second product to keep fp32-grade accuracy."""
    rng = np.random.default_rng(21)
    B = 600
    base = rng.standard_normal(64)
    U = base + 0.05 * rng.standard_normal((B, 64)); U /= np.linalg.norm(U, axis=1, keepdims=True)
    I = base + 0.05 * rng.standard_normal((B, 64)); I /= np.linalg.norm(I, axis=1, keepdims=True)
    loss, dU, dI = _run(U, I, 2)
    l64, rU, rI = O.in_batch_bpr_loss(U, I)
    assert abs(loss - float(l64)) <= 1e-6
    assert rel_l2(dU, rU) <= 1e-5 and rel_l2(dI, rI) <= 1e-5


def test_inbatch_tf32_fast_mode_stated_bound():
    """mode 1 = one TF32 pass per product: operands rounded to 10 mantissa bits (stated fast mode, not the parity mode)."""
    rng = np.random.default_rng(11)
    U, I = _unit(rng, 1024, 64), _unit(rng, 1024, 64)
    loss, dU, dI = _run(U, I, 1)
    l64, rU, rI = O.in_batch_bpr_loss(U, I)
    assert abs(loss - float(l64)) <= 1e-4
    assert rel_l2(dU, rU) <= 3e-3 and rel_l2(dI, rI) <= 3e-3


def test_inbatch_forward_only_and_deterministic():
    rng = np.random.default_rng(5)
    U, I = _unit(rng, 1500, 64), _unit(rng, 1500, 64)
    l_a, dU_a, dI_a = _run(U, I, 2)
    l_b, dU_b, dI_b = _run(U, I, 2)
    assert l_a == l_b and np.array_equal(dU_a, dU_b) and np.array_equal(dI_a, dI_b)
    with torch.no_grad():
        l_f, _, _ = _run(U, I, 2, grad=False)
    assert l_f == l_a


def test_inbatch_tc_rejects_other_widths():
    from recommendit_b200 import RB200Error
    rng = np.random.default_rng(0)
    with pytest.raises(RB200Error, match="D=64"):
        _run(_unit(rng, 16, 32), _unit(rng, 16, 32), 2)


def test_inbatch_full_batch_properties():
    """BASELINE batch (8192): too big for the fp64 oracle in seconds → size-independent properties.
    Σ_i dU_i·U_i + Σ_j dI_j·I_j = 0 does not hold in general, but the gradient of a function of S = U·Iᵀ satisfies
    Uᵀ·dU = dIᵀ·I (both equal Uᵀ·(G − diag(r))·I); and the two kernels must agree."""
    rng = np.random.default_rng(8192)
    B = 8192
    U, I = _unit(rng, B, 64).astype(np.float32), _unit(rng, B, 64).astype(np.float32)
    loss, dU, dI = _run(U, I, 2)
    loss0, dU0, dI0 = _run(U, I, 0)
    assert abs(loss - loss0) <= 1e-6
    assert rel_l2(dU, dU0) <= 1e-5 and rel_l2(dI, dI0) <= 1e-5
    lhs = U.astype(np.float64).T @ dU.astype(np.float64)
    rhs = dI.astype(np.float64).T @ I.astype(np.float64)
    assert rel_l2(lhs, rhs) <= 1e-5


def test_shared_memory_operand_kernel_stays_selectable():
    """RB200_INBATCH_TS=0 selects round 1's shared-memory-operand kernel (thread-staged Y tiles): the parity cases of this file
    again on it (the selector is read once per process, hence the subprocess)."""
    import os
    import subprocess
    import sys
    if os.environ.get("RB200_INBATCH_TS") == "0":
        pytest.skip("already inside the RB200_INBATCH_TS=0 run")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "tests/test_gpu_inbatch_tc.py", "-k", "not selectable"],
                       cwd=root, env=dict(os.environ, RB200_INBATCH_TS="0"), capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
