"""The threshold-pruned round algorithm of rb200_flat_search, restated on the CPU, against the plain exhaustive search — CPU."""
import numpy as np

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
