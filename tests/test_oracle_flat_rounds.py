"""The threshold-pruned round algorithm of rb200_flat_search, restated on the CPU, against the plain exhaustive search — CPU."""
import numpy as np
import pytest

from oracle import flat_rounds_oracle as FR
from oracle import ivf_oracle as V


def _exact(q, x, k):
    s = (q @ x.T).astype(np.float32)
    out_s, out_i = [], []
    for i in range(q.shape[0]):
        a, b = FR._topk_rows(s[i], np.arange(x.shape[0], dtype=np.int64), k)
        out_s.append(a); out_i.append(b)
    return np.stack(out_s), np.stack(out_i)


def test_pruned_rounds_equal_exhaustive_search_on_random_rows():
    rng = np.random.default_rng(0)
    x = V.normalize_rows(rng.standard_normal((6000, 16)).astype(np.float32))
    q = V.normalize_rows(rng.standard_normal((7, 16)).astype(np.float32))
    s, i, of = FR.flat_search_rounds(q, x, 50, prefix=256)
    es, ei = _exact(q, x, 50)
    assert not of and np.array_equal(i, ei) and np.array_equal(s, es)


def test_both_growth_regimes_are_exact():
    """more than 128 queries: rounds ×4, lists of 7·k; at most 128: rounds ×8, lists of 16·k (csrc/ivf.cu flat_growth / flat_cap)"""
    assert FR.constants_for(128) == (8, 16) and FR.constants_for(129) == (4, 7)
    rng = np.random.default_rng(7)
    x = V.normalize_rows(rng.standard_normal((9000, 16)).astype(np.float32))
    for nq in (5, 130):
        q = V.normalize_rows(rng.standard_normal((nq, 16)).astype(np.float32))
        s, i, of = FR.flat_search_rounds(q, x, 30, prefix=128)
        es, ei = _exact(q, x, 30)
        assert not of and np.array_equal(i, ei) and np.array_equal(s, es)


def test_exact_ties_go_to_the_earlier_row_across_rounds():
    rng = np.random.default_rng(1)
    x = V.normalize_rows(rng.standard_normal((3000, 8)).astype(np.float32))
    x[2500:2600] = x[100:200]                        # duplicates of early rows arrive in a later round
    q = x[100:110].copy()
    s, i, of = FR.flat_search_rounds(q, x, 20, prefix=128)
    es, ei = _exact(q, x, 20)
    assert not of and np.array_equal(i, ei) and np.array_equal(s, es)
    assert (i[:, 0] == np.arange(100, 110)).all()    # the original, not its later copy, wins the tie


def test_adversarial_row_order_overflows_and_is_redone_exactly():
    rng = np.random.default_rng(2)
    e = V.normalize_rows(rng.standard_normal((1, 8)).astype(np.float32))
    t = np.linspace(-1, 1, 4000, dtype=np.float32)[:, None]
    x = V.normalize_rows((t * e + 0.01 * rng.standard_normal((4000, 8))).astype(np.float32))
    x = x[np.argsort(x @ e[0], kind="stable")]       # every later row beats everything seen so far
    s, i, of = FR.flat_search_rounds(e, x, 10, prefix=64, cap=64)
    es, ei = _exact(e, x, 10)
    assert of and np.array_equal(i, ei) and np.array_equal(s, es)


def test_expected_survivors_per_round_are_about_growth_minus_one_times_k():
    """why the survivor list can be small: with exchangeable rows a round over [seen, 4·seen) leaves ≈ 3·k survivors per query"""
    rng = np.random.default_rng(3)
    x = V.normalize_rows(rng.standard_normal((32768, 8)).astype(np.float32))
    q = V.normalize_rows(rng.standard_normal((16, 8)).astype(np.float32))
    k, seen = 100, 8192
    s = q @ x.T
    thr = -np.sort(-s[:, :seen], axis=1)[:, k - 1]
    surv = (s[:, seen:] > thr[:, None]).sum(1)
    assert 2.0 * k < surv.mean() < 4.0 * k


@pytest.mark.parametrize("fmt", ["bf16", "tf32", "stream"])
@pytest.mark.parametrize("scale", ["unit", "wild"])
def test_filter_margin_never_loses_a_row_that_beats_the_threshold(fmt, scale):
    """The one-pass filters of the pruned rounds (bf16 / tf32 products + a margin in units of ||q||·max||x||): every score above its
    query's threshold in fp32 arithmetic passes the filter, whatever the scale of rows and queries, and the filter lets through only
    a little more than that (the survivors are re-scored in fp32, so extra ones cost time, not correctness)."""
    rng = np.random.default_rng({"bf16": 1, "tf32": 2, "stream": 3}[fmt])
    n, nq, d, k = 6000, 24, 64, 100
    x = V.normalize_rows(rng.standard_normal((n, d)).astype(np.float32))
    q = V.normalize_rows(rng.standard_normal((nq, d)).astype(np.float32))
    if scale == "wild":
        x = (x * np.exp(rng.uniform(np.log(0.05), np.log(20.0), (n, 1)))).astype(np.float32)
        q = (q * np.exp(rng.uniform(np.log(0.1), np.log(8.0), (nq, 1)))).astype(np.float32)
    s = (q.astype(np.float64) @ x.astype(np.float64).T)
    s32 = (q @ x.T).astype(np.float32)
    thr = -np.sort(-s32, axis=1)[:, k - 1]                       # each query's k-th best score: the hardest threshold to keep exact
    keep = FR.filter_keep(q, x, thr, fmt)
    must = s32 > thr[:, None]
    assert not (must & ~keep).any()                              # no winner lost
    # scores a whole margin below the threshold never pass (the filter is a filter)
    _, _, cm, _ = FR.FILTER[fmt]
    bound = 2.6 * cm * np.linalg.norm(q, axis=1)[:, None] * np.linalg.norm(x, axis=1).max()
    assert not (keep & (s < thr[:, None] - bound)).any()
    if scale == "unit":
        assert keep.sum() <= must.sum() + nq * (40 if fmt == "bf16" else 12)      # a few per cent of k extra per query
