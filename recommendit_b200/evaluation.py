"""Retrieval-quality harness: the reference's ranking metrics, batched over queries on whatever device
the id tensors live on, so that "IVF ids vs exact ids" and "trained towers vs held-out positives" are one call each.

Definitions follow ``src/evaluation/metrics.py`` of the reference (binary relevance):
  * ``recall_at_k``     :72-88   hits in the top k / |relevant|            (0 when nothing is relevant)
  * ``precision_at_k``  :90-101  hits in the top k / k
  * ``ndcg_at_k``       :20-69   DCG with 1/log2(rank+1) over the top k / ideal DCG of min(|relevant|, k) hits
  * ``mrr``             :103-119 1 / rank of the first relevant item (0 if none)
Inputs: ``recommended`` int64 [nq, n] ranked best first, padded with -1; ``relevant`` int64 [nq, r] padded with -1 (each row a
set: no duplicates).  These are index arithmetic on small tensors (analysis, not the hot path); tests pin them to the
reference's own functions through ``tests/golden/metrics.npz``.
"""
from __future__ import annotations

from typing import Dict

import torch


def _hits(recommended: torch.Tensor, relevant: torch.Tensor, k: int) -> torch.Tensor:
    """bool [nq, min(k, n)]: is the item at each rank relevant (padding never is)"""
    top = recommended[:, :k]
    rel_sorted, _ = torch.sort(relevant, dim=1)
    pos = torch.searchsorted(rel_sorted, top.contiguous()).clamp_(max=max(relevant.shape[1] - 1, 0))
    if relevant.shape[1] == 0:
        return torch.zeros_like(top, dtype=torch.bool)
    return (torch.gather(rel_sorted, 1, pos) == top) & (top >= 0)


def recall_at_k(recommended: torch.Tensor, relevant: torch.Tensor, k: int) -> torch.Tensor:
    n_rel = (relevant >= 0).sum(1)
    hits = _hits(recommended, relevant, k).sum(1)
    return torch.where(n_rel > 0, hits.double() / n_rel.clamp(min=1).double(), torch.zeros_like(hits, dtype=torch.float64))


def precision_at_k(recommended: torch.Tensor, relevant: torch.Tensor, k: int) -> torch.Tensor:
    if k == 0:
        return torch.zeros(recommended.shape[0], dtype=torch.float64, device=recommended.device)
    return _hits(recommended, relevant, k).sum(1).double() / k


def ndcg_at_k(recommended: torch.Tensor, relevant: torch.Tensor, k: int) -> torch.Tensor:
    h = _hits(recommended, relevant, k).double()
    disc = 1.0 / torch.log2(torch.arange(2, h.shape[1] + 2, device=h.device, dtype=torch.float64))
    dcg = (h * disc).sum(1)
    n_ideal = (relevant >= 0).sum(1).clamp(max=k)
    full = 1.0 / torch.log2(torch.arange(2, k + 2, device=h.device, dtype=torch.float64))
    cum = torch.cat([torch.zeros(1, device=h.device, dtype=torch.float64), torch.cumsum(full, 0)])
    idcg = cum[n_ideal]
    return torch.where(idcg > 0, dcg / idcg.clamp(min=1e-300), torch.zeros_like(dcg))


def mrr(recommended: torch.Tensor, relevant: torch.Tensor) -> torch.Tensor:
    h = _hits(recommended, relevant, recommended.shape[1])
    any_hit = h.any(1)
    first = torch.argmax(h.to(torch.int8), dim=1)
    return torch.where(any_hit, 1.0 / (first.double() + 1.0), torch.zeros(h.shape[0], dtype=torch.float64, device=h.device))


def retrieval_report(approx_ids: torch.Tensor, exact_ids: torch.Tensor, ks=(10, 100, 500)) -> Dict[str, float]:
    """Mean Recall@K / NDCG@K of an approximate top-k (IVF) with the exact top-K as the relevant set, per cutoff K."""
    out = {}
    for k in ks:
        if k > exact_ids.shape[1]:
            continue
        rel = exact_ids[:, :k]
        out[f"recall@{k}"] = float(recall_at_k(approx_ids, rel, k).mean())
        out[f"ndcg@{k}"] = float(ndcg_at_k(approx_ids, rel, k).mean())
    return out
