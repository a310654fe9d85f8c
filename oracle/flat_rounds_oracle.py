"""CPU restatement of the threshold-pruned exhaustive search (rb200_flat_search's round loop, csrc/ivf.cu + csrc/flat_scan_tc.cu)
— TEST INFRASTRUCTURE ONLY.  It restates the ALGORITHM, not the arithmetic: scores come from the same fp32 matmul as
``ivf_oracle.flat_search``; what is checked (tests/test_oracle_flat_rounds.py) is that pruning by the running k-th best score with a
strict comparison, row ranges growing by a factor of four (eight for at most 128 queries), a bounded survivor list and the
redo-on-overflow rule return exactly the
exhaustive top-k (faiss IndexFlatIP semantics: score descending, earlier row first on equal scores) for any row order.
"""
import numpy as np

PREFIX, GROWTH, CAP_K, CAP_MIN = 8192, 4, 7, 3584       # the constants of csrc/ivf.cu (more than 128 queries)
GROWTH_SMALL, CAP_K_SMALL = 8, 16                       # at most 128 queries: flat_growth / flat_cap in csrc/ivf.cu


def constants_for(nq):
    """(growth, cap_k) rb200_flat_search uses for a batch of nq queries"""
    return (GROWTH_SMALL, CAP_K_SMALL) if nq <= 128 else (GROWTH, CAP_K)


def _topk_rows(scores, rows, k):
    """(score desc, row asc) top-k of one query's candidates"""
    order = np.lexsort((rows, -scores.astype(np.float64)))[:k]
    return scores[order], rows[order]


def flat_search_rounds(q, x, k, prefix=PREFIX, growth=None, cap=None):
    """→ (scores [nq, k], rows [nq, k], overflowed: bool).  -FLT_MAX / -1 padded when k > rows."""
    q, x = np.asarray(q, np.float32), np.asarray(x, np.float32)
    nq, n = q.shape[0], x.shape[0]
    g_auto, cap_k = constants_for(nq)
    if growth is None:
        growth = g_auto
    if cap is None:
        cap = (max(cap_k * k, CAP_MIN) + 511) // 512 * 512
    out_s = np.full((nq, k), -np.finfo(np.float32).max, np.float32)
    out_i = np.full((nq, k), -1, np.int64)
    n0 = min(n, prefix)
    s0 = q @ x[:n0].T
    best = []
    for i in range(nq):
        best.append(_topk_rows(s0[i], np.arange(n0, dtype=np.int64), k))
    overflow = False
    seen = n0
    while seen < n:
        upto = min(n, seen * growth)
        sc = q @ x[seen:upto].T
        for i in range(nq):
            bs, bi = best[i]
            thr = bs[k - 1] if len(bs) >= k else -np.inf           # k-th best so far
            surv = np.nonzero(sc[i] > thr)[0]                       # strict: an equal score loses to the earlier row
            if len(surv) > cap:
                overflow = True                                     # the kernel drops what does not fit and flags the search
                surv = surv[:cap]
            cs = np.concatenate([bs, sc[i][surv]])
            ci = np.concatenate([bi, surv.astype(np.int64) + seen])
            best[i] = _topk_rows(cs, ci, k)
        seen = upto
    if overflow:                                                    # redone on the chunked path: exact for any row order
        s_all = q @ x.T
        best = [_topk_rows(s_all[i], np.arange(n, dtype=np.int64), k) for i in range(nq)]
    for i in range(nq):
        m = len(best[i][0])
        out_s[i, :m], out_i[i, :m] = best[i]
    return out_s, out_i, overflow


# ---------------------------------------------------------------------------------------------------------------- #
# The one-pass FILTER of the pruned rounds (csrc/flat_filter_tc.cu, csrc/flat_stream_tc.cu), restated: operands rounded to bf16
# (round-to-nearest-even, both rows and queries) or to tf32 (rows truncated / queries rounded), products accumulated in fp32 and a
# score kept when  S~ + margin_q * max||x|| - thr_q > 0  with max||x|| over the rows of a CTA (256 rows) or of a tile (128 rows).
# What the tests check on the CPU: the kept set is a SUPERSET of {S > thr} for any scale of rows and queries (no true winner can be
# lost), and it is not much larger.
# ---------------------------------------------------------------------------------------------------------------- #
def _bf16_rn(a):
    """float32 -> the nearest bfloat16 (ties to even), returned as float32"""
    u = np.asarray(a, np.float32).view(np.uint32).astype(np.uint64)
    r = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return r.astype(np.uint32).view(np.float32)


def _tf32_rn(a):
    """float32 -> tf32 by adding half an ulp to the magnitude and truncating (umma::tf32_hi)"""
    u = np.asarray(a, np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)


def _tf32_trunc(a):
    """what kind::tf32 does to a raw fp32 operand: the low 13 mantissa bits are ignored"""
    return (np.asarray(a, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


FILTER = {                      # format -> (row rounding, query rounding, margin constant C: margin_q = C * ||q||, rows per norm block)
    "bf16": (_bf16_rn, _bf16_rn, 1.5 / 256.0, 256),             # flat_filter_tc_kernel<true>   (more than 128 queries)
    "tf32": (_tf32_rn, _tf32_rn, 1.5 / 1024.0, 256),            # flat_filter_tc_kernel<false>
    "stream": (_tf32_trunc, _tf32_rn, 2.0 / 1024.0, 128),       # flat_stream_tc_kernel (raw rows by TMA, margin 2^-9, per-tile norm)
}


def filter_keep(q, x, thr, fmt="bf16"):
    """-> boolean [nq, n]: the scores the filter lets through (they are re-scored in fp32 afterwards)"""
    rx, rq, cm, blk = FILTER[fmt]
    q, x = np.asarray(q, np.float32), np.asarray(x, np.float32)
    s_approx = (rq(q).astype(np.float32) @ rx(x).astype(np.float32).T).astype(np.float32)
    qn = np.linalg.norm(q.astype(np.float64), axis=1) * 1.0001
    xn = np.linalg.norm(x.astype(np.float64), axis=1) * 1.0001
    n = x.shape[0]
    nmax = np.empty(n)
    for b in range(0, n, blk):
        nmax[b:b + blk] = xn[b:b + blk].max()
    margin = (cm * qn)[:, None] * nmax[None, :]
    return s_approx.astype(np.float64) + margin - np.asarray(thr, np.float64)[:, None] > 0.0
