"""CPU stand-in for recommendit_b200.sharded.CudaOps built on the NumPy oracle — TEST infrastructure.
Lets the host-side exchange logic of the sharded trainer (routing, all-to-all splits, gradient scaling, all-reduces)
run under gloo on CPU.  Never used by the product."""
import math

import numpy as np
import torch

from oracle import ivf_oracle as V
from oracle import two_tower_oracle as O
from recommendit_b200._lib import OptState


def _np(t):
    return t.detach().cpu().numpy()


class OracleOps:
    def route(self, user_ids, item_ids, world, nu_by_rank):
        from recommendit_b200.sharded import route_reference
        return route_reference(user_ids, item_ids, world, nu_by_rank)

    def gather_rows(self, table, rows):
        out = torch.zeros(rows.numel(), table.shape[1], dtype=torch.float32)
        ok = (rows >= 0) & (rows < table.shape[0])
        out[ok] = table[rows[ok]]
        return out

    def towers_fwd(self, jobs, D, H, drop_p, seed, offset):
        for j in jobs:
            din = D + (0 if j.get("extra") is None else j["extra"].shape[1])
            y, c = O.tower_forward(_np(j["table"]), _np(j["ids"]), None if j.get("extra") is None else _np(j["extra"]),
                                   _np(j["W1"]).reshape(H, din), _np(j["b1"]), _np(j["W2"]).reshape(D, H), _np(j["b2"]))
            j["out"].copy_(torch.from_numpy(y)); j["hid"].copy_(torch.from_numpy(c.h)); j["denom"].copy_(torch.from_numpy(c.denom[:, 0]))
            j["_cache"] = c

    def bpr_pair(self, u, p, n, grad_scale):
        loss, du, dp, dn = O.bpr_loss(_np(u), _np(p), _np(n))
        s = np.float32(grad_scale)
        return (torch.tensor([float(loss)], dtype=torch.float32), torch.from_numpy(du * s), torch.from_numpy(dp * s),
                torch.from_numpy(dn * s))

    def towers_bwd(self, jobs, D, H, drop_p, grads_out):
        acc = None
        for j in jobs:
            dW1, db1, dW2, db2, dr = O.tower_backward(j["_cache"], _np(j["dY"]))
            flat = np.concatenate([dW1.ravel(), db1.ravel(), dW2.ravel(), db2.ravel()])
            acc = flat if acc is None else acc + flat
            j["dRows"].copy_(torch.from_numpy(np.ascontiguousarray(dr)))
        grads_out.copy_(torch.from_numpy(acc.astype(np.float32)))

    def scatter_rows(self, ids, rows, n_rows, padding_row):
        B, D = rows.shape
        cap = max(B, 1)
        ids_np, rows_np = _np(ids), _np(rows)
        keep = (ids_np != padding_row) & (ids_np >= 0) & (ids_np < n_rows)
        uniq = np.unique(ids_np[keep])
        ug = np.zeros((cap, D), np.float32)
        for s, r in enumerate(uniq):
            for i in np.nonzero(ids_np == r)[0]:          # ascending sample order
                ug[s] += rows_np[i]
        u = np.zeros(cap, np.int64); u[: len(uniq)] = uniq
        return torch.from_numpy(u), torch.from_numpy(ug), torch.tensor([len(uniq)], dtype=torch.int32)

    def read_opt(self, opt):
        return OptState.from_buffer_copy(bytes(opt.cpu().numpy().tobytes()))

    def write_opt(self, opt, st):
        opt.copy_(torch.frombuffer(bytearray(bytes(st)), dtype=torch.uint8))

    def begin_step(self, opt):
        st = self.read_opt(opt)
        st.step += 1
        st.step_size = st.lr / (1 - st.beta1 ** st.step)
        st.bias_corr2_sqrt = math.sqrt(1 - st.beta2 ** st.step)
        st.sumsq = 0.0
        self.write_opt(opt, st)

    def grad_norm_clip(self, opt):
        st = self.read_opt(opt)
        st.total_norm = math.sqrt(st.sumsq)
        st.clip_coef = min(1.0, st.max_norm / (st.total_norm + 1e-6))
        self.write_opt(opt, st)

    def sumsq(self, opt, segs):
        st = self.read_opt(opt)
        for t, cnt, rl in segs:
            n = t.numel() if cnt is None else int(cnt.item()) * rl
            st.sumsq += float((_np(t).reshape(-1)[:n].astype(np.float64) ** 2).sum())
        self.write_opt(opt, st)

    def _adam(self, w, g, m, v, st):
        nw, nm, nv = O.adam_step(_np(w), _np(g) * np.float32(st.clip_coef), _np(m), _np(v), int(st.step), st.lr, st.beta1, st.beta2,
                                 st.eps, st.weight_decay)
        w.copy_(torch.from_numpy(nw)); m.copy_(torch.from_numpy(nm)); v.copy_(torch.from_numpy(nv))

    def adam_dense(self, w, g, m, v, opt):
        self._adam(w, g, m, v, self.read_opt(opt))

    def adam_rows(self, w, m, v, uniq, ug, nu, opt):
        n = int(nu.item())
        rows = uniq[:n]
        ww, mm, vv = w[rows].clone(), m[rows].clone(), v[rows].clone()
        self._adam(ww, ug[:n], mm, vv, self.read_opt(opt))
        w[rows], m[rows], v[rows] = ww, mm, vv

    def adam_table_dense(self, w, m, v, uniq, ug, nu, slot, opt):
        n = int(nu.item())
        g = torch.zeros_like(w)
        g[uniq[:n]] = ug[:n]
        self._adam(w, g, m, v, self.read_opt(opt))


def flat_search_cpu(q, db, k, id_base):
    s, i = V.flat_search(_np(q), _np(db), k)
    return torch.from_numpy(s), torch.from_numpy(np.where(i >= 0, i + id_base, -1))


def topk_merge_cpu(scores, ids):
    parts, nq, k = scores.shape
    s = _np(scores).transpose(1, 0, 2).reshape(nq, parts * k)
    i = _np(ids).transpose(1, 0, 2).reshape(nq, parts * k)
    top = np.argsort(-s, axis=1, kind="stable")[:, :k]
    return torch.from_numpy(np.take_along_axis(s, top, 1)), torch.from_numpy(np.take_along_axis(i, top, 1))
