"""Golden vectors for recommendit_b200/evaluation.py from the reference's own metric functions (src/evaluation/metrics.py).
Run in the build container (needs /root/reference):  python tests/golden/make_metrics_golden.py"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, "/root/reference")
from src.evaluation.metrics import mrr, ndcg_at_k, precision_at_k, recall_at_k  # noqa: E402

rng = np.random.default_rng(5)
nq, n, r = 40, 30, 12
rec = np.full((nq, n), -1, np.int64)
rel = np.full((nq, r), -1, np.int64)
for q in range(nq):
    m = int(rng.integers(0, n + 1))                       # ragged recommendation lists (incl. empty)
    rec[q, :m] = rng.permutation(60)[:m]
    t = int(rng.integers(0, r + 1))                       # ragged relevant sets (incl. empty)
    rel[q, :t] = rng.permutation(60)[:t]
out = {"rec": rec, "rel": rel}
for k in (1, 5, 10, 30):
    for name, fn in (("recall", recall_at_k), ("precision", precision_at_k), ("ndcg", ndcg_at_k)):
        out[f"{name}@{k}"] = np.array([fn([int(x) for x in rec[q] if x >= 0], [int(x) for x in rel[q] if x >= 0], k) for q in range(nq)])
out["mrr"] = np.array([mrr([int(x) for x in rec[q] if x >= 0], [int(x) for x in rel[q] if x >= 0]) for q in range(nq)])
np.savez(Path(__file__).parent / "metrics.npz", **out)
print("wrote metrics.npz")
