"""CPU oracle for the two-tower BPR training step.  TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement of the arithmetic the reference delegates to
PyTorch on its two-tower hot path.  It is the *checker* for the CUDA kernels in
``recommendit_b200/csrc``; nothing in the product package may import it.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs are allowed to use it.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the unmodified
reference (``/root/reference/src/models/two_tower.py``) in the build container,
runs it under fixed seeds and stores inputs + outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function here against those
fixtures.

Reference lines each function follows (paths relative to /root/reference):

* ``tower_forward``            src/models/two_tower.py:39-42 (user), :68-72 (item)
* ``tower_backward``           autograd of the above (train_embeddings.py:190)
* ``embedding_dense_backward`` autograd of nn.Embedding(padding_idx=0), two_tower.py:27,54
* ``bpr_loss``                 src/models/two_tower.py:117-130
* ``in_batch_bpr_loss_loop``   src/models/two_tower.py:132-160 (the literal loop)
* ``in_batch_bpr_loss``        closed form of the same
* ``clip_grad_norm``           src/training/train_embeddings.py:191
* ``adam_step``                src/training/train_embeddings.py:160,192 (torch.optim.Adam,
                               coupled weight_decay, bias correction, amsgrad=False)
* ``train_step``               src/training/train_embeddings.py:183-192

All functions take a ``dtype`` implicitly from their inputs: feed float32 for a
like-for-like check, float64 for a tighter "true value" reference.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

N_GENRES = 18  # two_tower.py:16

NORMALIZE_EPS = 1e-12  # F.normalize default eps (two_tower.py:42,72)


# --------------------------------------------------------------------------- #
# towers
# --------------------------------------------------------------------------- #
@dataclass
class TowerCache:
    ids: np.ndarray
    x: np.ndarray          # [B, Din] MLP input (embedding row ++ extra)
    h: np.ndarray          # [B, H]   post-ReLU, post-dropout hidden
    keep: np.ndarray       # [B, H]   multiplicative factor applied after ReLU (mask/(1-p)), 1 in eval
    pre: np.ndarray        # [B, D]   pre-normalisation output
    denom: np.ndarray      # [B, 1]   max(||pre||, eps)
    y: np.ndarray          # [B, D]   normalised output
    W1: np.ndarray
    W2: np.ndarray
    D: int


def tower_forward(table, ids, extra, W1, b1, W2, b2, drop_mask=None, drop_p: float = 0.0):
    """normalize(W2 · drop(relu(W1 · [E[ids]; extra] + b1)) + b2)   (two_tower.py:39-42, 68-72).

    ``drop_mask`` is a {0,1} array [B,H] of kept units (None ⇒ eval / p=0).
    Inverted-dropout scaling 1/(1-p) as torch.nn.Dropout.
    """
    ids = np.asarray(ids, dtype=np.int64)
    e = table[ids]                                            # nn.Embedding lookup
    x = e if extra is None else np.concatenate([e, extra.astype(e.dtype)], axis=-1)  # torch.cat (:70)
    a = x @ W1.T + b1                                         # Linear: x·Wᵀ + b, W is [out, in]
    r = np.maximum(a, 0)
    if drop_mask is not None and drop_p > 0.0:
        keep = drop_mask.astype(r.dtype) / r.dtype.type(1.0 - drop_p)
    else:
        keep = np.ones_like(r)
    h = r * keep
    pre = h @ W2.T + b2
    nrm = np.sqrt((pre * pre).sum(-1, keepdims=True))
    denom = np.maximum(nrm, pre.dtype.type(NORMALIZE_EPS))    # clamp_min(eps)
    y = pre / denom
    return y, TowerCache(ids, x, h, keep, pre, denom, y, W1, W2, table.shape[1])


def tower_backward(c: TowerCache, dy):
    """Backward of ``tower_forward``: returns dW1, db1, dW2, db2, dRows[B,D].

    normalize backward: for ||pre|| > eps, d pre = (g − y (y·g)) / ||pre||; when the clamp is
    active the output is pre/eps (a pure scale) so d pre = g/eps.
    """
    g = dy
    nrm = np.sqrt((c.pre * c.pre).sum(-1, keepdims=True))
    clamped = nrm < c.pre.dtype.type(NORMALIZE_EPS)
    dpre = (g - c.y * (c.y * g).sum(-1, keepdims=True)) / c.denom
    dpre = np.where(clamped, g / c.denom, dpre)
    dW2 = dpre.T @ c.h
    db2 = dpre.sum(0)
    dh = dpre @ c.W2
    da = dh * c.keep * (c.h > 0)        # relu'(a)=1 iff a>0; h>0 ⇔ a>0 and kept
    # units that were kept but had a<=0 have h==0 → zero grad, same as autograd.
    dW1 = da.T @ c.x
    db1 = da.sum(0)
    dx = da @ c.W1
    return dW1, db1, dW2, db2, dx[:, : c.D]


def embedding_dense_backward(ids, drows, n_rows: int, padding_idx: int = 0):
    """Dense [n_rows, D] gradient of an nn.Embedding(padding_idx=0) lookup: duplicates are summed,
    rows equal to padding_idx receive nothing (two_tower.py:27,54)."""
    ids = np.asarray(ids, dtype=np.int64)
    out = np.zeros((n_rows, drows.shape[1]), dtype=drows.dtype)
    sel = ids != padding_idx
    np.add.at(out, ids[sel], drows[sel])
    return out


# --------------------------------------------------------------------------- #
# losses
# --------------------------------------------------------------------------- #
def _softplus(x):
    # -logsigmoid(d) = softplus(-d); torch: min(0,d) - log1p(exp(-|d|)) negated
    return np.maximum(x, 0) + np.log1p(np.exp(-np.abs(x)))


def _sigmoid(x):
    e = np.exp(-np.abs(x))
    return np.where(x >= 0, 1 / (1 + e), e / (1 + e))


def bpr_loss(u, p, n):
    """-logsigmoid((u·p) − (u·n)).mean()  (two_tower.py:127-129) and its gradients."""
    B = u.shape[0]
    pos = (u * p).sum(-1)
    neg = (u * n).sum(-1)
    d = pos - neg
    loss = _softplus(-d).mean()
    gd = (-_sigmoid(-d) / B)[:, None].astype(u.dtype)      # dL/dd
    du = gd * (p - n)
    dp = gd * u
    dn = -gd * u
    return loss.astype(u.dtype), du, dp, dn


def in_batch_bpr_loss_loop(U, I):
    """Literal restatement of the reference's Python loop (two_tower.py:143-160); small B only."""
    S = U @ I.T
    B = U.shape[0]
    pos = np.diag(S)
    loss = U.dtype.type(0)
    for i in range(B):
        mask = np.ones(B, dtype=bool)
        mask[i] = False
        margins = pos[i] - S[i][mask]
        loss = loss + _softplus(-margins).mean()
    return (loss / B).astype(U.dtype)


def in_batch_bpr_loss(U, I):
    """Closed form  Σ_{i≠j} softplus(S_ij − S_ii) / (B(B−1))  and gradients dU, dI."""
    B = U.shape[0]
    S = U @ I.T
    dg = np.diag(S).copy()
    M = S - dg[:, None]
    off = ~np.eye(B, dtype=bool)
    denom = U.dtype.type(B * (B - 1))
    loss = (_softplus(M) * off).sum() / denom
    G = _sigmoid(M) * off / denom                # dL/dS_ij for j≠i
    G[np.arange(B), np.arange(B)] = -G.sum(1)    # dL/dS_ii
    G = G.astype(U.dtype)
    return loss.astype(U.dtype), G @ I, G.T @ U


# --------------------------------------------------------------------------- #
# clip + Adam
# --------------------------------------------------------------------------- #
def clip_grad_norm(grads: List[np.ndarray], max_norm: float = 1.0):
    """torch.nn.utils.clip_grad_norm_: total = ‖[‖g_p‖₂]_p‖₂ ; coef = min(1, max_norm/(total+1e-6))."""
    dt = grads[0].dtype
    norms = np.array([np.sqrt((g.astype(dt) ** 2).sum()) for g in grads], dtype=dt)
    total = np.sqrt((norms ** 2).sum())
    coef = min(1.0, float(max_norm) / (float(total) + 1e-6))
    return dt.type(coef), dt.type(total)


def adam_step(w, g, m, v, step: int, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-5):
    """One torch.optim.Adam update (non-amsgrad, coupled L2): returns (w, m, v) new arrays.
    ``step`` is the 1-based step count *after* increment, as torch uses it for bias correction."""
    dt = w.dtype.type
    g = g + dt(weight_decay) * w
    m = m + (g - m) * dt(1 - beta1)                 # torch: exp_avg.lerp_(grad, 1-beta1)
    v = v * dt(beta2) + dt(1 - beta2) * g * g       # exp_avg_sq.mul_(b2).addcmul_(g,g,1-b2)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = dt(lr / bc1)
    denom = np.sqrt(v) / dt(np.sqrt(bc2)) + dt(eps)
    w = w - step_size * (m / denom)
    return w, m, v


# --------------------------------------------------------------------------- #
# the whole step (train_embeddings.py:183-192)
# --------------------------------------------------------------------------- #
PARAM_KEYS = (
    "user_tower.embedding.weight", "user_tower.mlp.0.weight", "user_tower.mlp.0.bias",
    "user_tower.mlp.3.weight", "user_tower.mlp.3.bias",
    "item_tower.embedding.weight", "item_tower.mlp.0.weight", "item_tower.mlp.0.bias",
    "item_tower.mlp.3.weight", "item_tower.mlp.3.bias",
)


def _tower_args(P, which):
    t = which + "_tower."
    return (P[t + "embedding.weight"], P[t + "mlp.0.weight"], P[t + "mlp.0.bias"],
            P[t + "mlp.3.weight"], P[t + "mlp.3.bias"])


def loss_and_grads(P: Dict[str, np.ndarray], user_ids, pos_ids, pos_g, neg_ids, neg_g,
                   masks=None, drop_p: float = 0.0, in_batch: bool = False):
    """Forward ×3 + bpr_loss + backward.  Returns (loss, grads dict keyed like PARAM_KEYS,
    (u, p, n) embeddings).  ``masks`` = optional (mu, mp, mn) dropout keep-masks."""
    mu, mp, mn = masks if masks is not None else (None, None, None)
    ut, uW1, ub1, uW2, ub2 = _tower_args(P, "user")
    it, iW1, ib1, iW2, ib2 = _tower_args(P, "item")
    u, cu = tower_forward(ut, user_ids, None, uW1, ub1, uW2, ub2, mu, drop_p)
    p, cp = tower_forward(it, pos_ids, pos_g, iW1, ib1, iW2, ib2, mp, drop_p)
    if in_batch:
        loss, du, dp = in_batch_bpr_loss(u, p)
        n, cn, dn = None, None, None
    else:
        n, cn = tower_forward(it, neg_ids, neg_g, iW1, ib1, iW2, ib2, mn, drop_p)
        loss, du, dp, dn = bpr_loss(u, p, n)
    G = {}
    dW1, db1, dW2, db2, dr = tower_backward(cu, du)
    G["user_tower.embedding.weight"] = embedding_dense_backward(user_ids, dr, ut.shape[0])
    G["user_tower.mlp.0.weight"], G["user_tower.mlp.0.bias"] = dW1, db1
    G["user_tower.mlp.3.weight"], G["user_tower.mlp.3.bias"] = dW2, db2
    dW1, db1, dW2, db2, dr = tower_backward(cp, dp)
    gi = embedding_dense_backward(pos_ids, dr, it.shape[0])
    if not in_batch:
        eW1, eb1, eW2, eb2, er = tower_backward(cn, dn)
        dW1, db1, dW2, db2 = dW1 + eW1, db1 + eb1, dW2 + eW2, db2 + eb2
        gi = gi + embedding_dense_backward(neg_ids, er, it.shape[0])
    G["item_tower.embedding.weight"] = gi
    G["item_tower.mlp.0.weight"], G["item_tower.mlp.0.bias"] = dW1, db1
    G["item_tower.mlp.3.weight"], G["item_tower.mlp.3.bias"] = dW2, db2
    return loss, G, (u, p, n)


@dataclass
class AdamState:
    step: int = 0
    m: Dict[str, np.ndarray] = field(default_factory=dict)
    v: Dict[str, np.ndarray] = field(default_factory=dict)


def train_step(P, S: AdamState, batch, lr=1e-3, weight_decay=1e-5, max_norm=1.0,
               masks=None, drop_p=0.0, in_batch=False, beta1=0.9, beta2=0.999, eps=1e-8):
    """One optimiser step exactly as train_embeddings.py:183-192: forward, loss, backward,
    clip_grad_norm_(1.0), Adam(wd).  Mutates P and S in place; returns (loss, grads-before-clip,
    clip coefficient)."""
    loss, G, _ = loss_and_grads(P, *batch, masks=masks, drop_p=drop_p, in_batch=in_batch)
    coef, total = clip_grad_norm([G[k] for k in PARAM_KEYS], max_norm)
    S.step += 1
    for k in PARAM_KEYS:
        if k not in S.m:
            S.m[k] = np.zeros_like(P[k])
            S.v[k] = np.zeros_like(P[k])
        P[k], S.m[k], S.v[k] = adam_step(P[k], G[k] * coef, S.m[k], S.v[k], S.step, lr,
                                         beta1, beta2, eps, weight_decay)
    return loss, G, coef


def init_params(n_users: int, n_items: int, embed_dim: int = 64, hidden_dim: int = 128,
                seed: int = 0, dtype=np.float32) -> Dict[str, np.ndarray]:
    """Random parameters with the reference's *distributions* (two_tower.py:27-37, 54-66;
    SURVEY Appendix A): xavier-uniform tables incl. row 0, U(±1/√fan_in) Linear weights/biases.
    (Not the same RNG stream as torch — parity runs copy parameters across instead.)"""
    rng = np.random.default_rng(seed)

    def U(shape, a):
        return rng.uniform(-a, a, size=shape).astype(dtype)

    P = {}
    for name, n, din in (("user", n_users, embed_dim), ("item", n_items, embed_dim + N_GENRES)):
        t = name + "_tower."
        P[t + "embedding.weight"] = U((n + 1, embed_dim), np.sqrt(6.0 / (n + 1 + embed_dim)))
        P[t + "mlp.0.weight"] = U((hidden_dim, din), 1 / np.sqrt(din))
        P[t + "mlp.0.bias"] = U((hidden_dim,), 1 / np.sqrt(din))
        P[t + "mlp.3.weight"] = U((embed_dim, hidden_dim), 1 / np.sqrt(hidden_dim))
        P[t + "mlp.3.bias"] = U((embed_dim,), 1 / np.sqrt(hidden_dim))
    return P
