"""tcgen05 (kind::tf32) score product against an fp64 reference: 3xTF32 must be fp32-grade, TF32 within its stated bound."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 128, 64), (256, 384, 64), (100, 70, 64), (4096, 4096, 64),
                                   (333, 129, 128), (64, 4096, 88), (1, 1, 4), (130, 250, 36)])
@pytest.mark.parametrize("mode", [2, 1])
def test_gemm_nt_matches_fp64(M, N, K, mode):
    import recommendit_b200 as R
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g)
    b = torch.randn(N, K, generator=g)
    a, b = torch.nn.functional.normalize(a, dim=-1), torch.nn.functional.normalize(b, dim=-1)
    ref = a.double() @ b.double().T
    out = R.scores_nt(a.cuda(), b.cuda(), mode=mode).cpu().double()
    err = (out - ref).abs().max().item()
    # unit-norm rows ⇒ |score| ≤ 1.  3xTF32: ~2^-21 per product term; TF32: 2^-11 per operand.
    assert err <= (2e-6 if mode == 2 else 2e-3), (M, N, K, mode, err)
    if mode == 2:
        fp32 = (a.cuda() @ b.cuda().T).cpu().double()      # cuBLAS fp32 for scale: we should be in the same class
        assert err <= 4 * max((fp32 - ref).abs().max().item(), 2e-7)


def test_gemm_nt_exact_on_representable_inputs():
    """Small integers are exact in tf32: the product must be exact in every mode (checks layouts, not rounding)."""
    import recommendit_b200 as R
    g = torch.Generator().manual_seed(0)
    a = torch.randint(-4, 5, (257, 72), generator=g).float()
    b = torch.randint(-4, 5, (190, 72), generator=g).float()
    ref = a @ b.T
    for mode in (1, 2):
        out = R.scores_nt(a.cuda(), b.cuda(), mode=mode).cpu()
        assert torch.equal(out, ref), mode

