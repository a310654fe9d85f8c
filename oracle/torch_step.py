"""CPU baseline arm: the reference's training-step body restated with the SAME library calls it makes.
TEST / BENCH INFRASTRUCTURE ONLY (bench.py ``cpu_baseline`` and ``--impl reference``; tests).

The reference's hot loop (src/training/train_embeddings.py:183-192) is Python that calls PyTorch:
``nn.Embedding`` → ``nn.Linear`` → ``ReLU`` → ``Dropout`` → ``nn.Linear`` → ``F.normalize`` three times
(src/models/two_tower.py:39-42, 68-72), ``bpr_loss`` (:127-129), ``zero_grad`` / ``backward`` /
``clip_grad_norm_(1.0)`` / ``torch.optim.Adam(lr, weight_decay=1e-5).step()`` / ``loss.item()``.
``/root/reference`` does not exist on the GPU box, so this module issues exactly those ATen calls through
``torch.nn.functional`` on parameters laid out like the reference's ``state_dict``.  It is "kind: port".
Pinned by tests/test_oracle_golden.py::test_torch_step_matches_reference_golden.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .two_tower_oracle import PARAM_KEYS


def make_params(P: Dict[str, np.ndarray]) -> Dict[str, torch.Tensor]:
    return {k: torch.tensor(np.asarray(P[k], dtype=np.float32), requires_grad=True) for k in PARAM_KEYS}


def make_optimizer(T: Dict[str, torch.Tensor], lr: float = 1e-3, weight_decay: float = 1e-5) -> torch.optim.Adam:
    return torch.optim.Adam([T[k] for k in PARAM_KEYS], lr=lr, weight_decay=weight_decay)   # train_embeddings.py:160


def tower(T, which: str, ids: torch.Tensor, extra, p: float, training: bool) -> torch.Tensor:
    t = which + "_tower."
    x = F.embedding(ids, T[t + "embedding.weight"], padding_idx=0)                 # two_tower.py:40 / :69
    if extra is not None:
        x = torch.cat([x, extra], dim=-1)                                          # :70
    x = F.linear(x, T[t + "mlp.0.weight"], T[t + "mlp.0.bias"])
    x = F.dropout(F.relu(x), p=p, training=training)
    x = F.linear(x, T[t + "mlp.3.weight"], T[t + "mlp.3.bias"])
    return F.normalize(x, p=2, dim=-1)                                             # :42 / :72


def step(T, opt, batch: Tuple[torch.Tensor, ...], dropout: float = 0.1, training: bool = True) -> float:
    """train_embeddings.py:183-194 on CPU tensors."""
    user_ids, pos_ids, pos_genres, neg_ids, neg_genres = batch
    user_emb = tower(T, "user", user_ids, None, dropout, training)
    pos_emb = tower(T, "item", pos_ids, pos_genres, dropout, training)
    neg_emb = tower(T, "item", neg_ids, neg_genres, dropout, training)
    pos_scores = (user_emb * pos_emb).sum(dim=-1)                                  # two_tower.py:127-129
    neg_scores = (user_emb * neg_emb).sum(dim=-1)
    loss = -F.logsigmoid(pos_scores - neg_scores).mean()
    opt.zero_grad()
    loss.backward()
    torch.nn.utils.clip_grad_norm_([T[k] for k in PARAM_KEYS], max_norm=1.0)
    opt.step()
    return loss.item()
