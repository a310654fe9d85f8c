"""GPU parity of the fused one-call training step (rb200_bpr_step) against the reference trajectory."""
import numpy as np
import pytest
import torch

from oracle import two_tower_oracle as O
from tests.parity import (batch_from_golden, dev, flat_mlp, grad_tol, masks_from_golden, model_from_golden,
                          params_from_golden, rel_l2)

pytestmark = pytest.mark.gpu
CASES = ["tt_small", "tt_dup", "tt_dropout", "tt_d128", "tt_drop64"]


def _trainer(model, g, **kw):
    import recommendit_b200 as R
    return R.FusedBPRTrainer(model, lr=float(g["lr"]), weight_decay=1e-5, max_norm=1.0, **kw)


@pytest.mark.parametrize("case", CASES)
def test_fused_step_matches_reference_trajectory(golden, case):
    g = golden(case)
    model = model_from_golden(g).train()
    tr = _trainer(model, g, use_cuda_graph=False)
    lr = float(g["lr"])
    for s in range(int(g["meta"][5])):
        pre = f"step{s}/"
        b = batch_from_golden(g, s)
        tr.load_packed(tr.pack_host(*b))
        masks = masks_from_golden(g, s)
        loss = tr.step(masks=None if masks is None else [torch.from_numpy(m) for m in masks]).item()
        assert abs(loss - float(g[pre + "loss"])) <= 1e-6 + 4e-5 * s, (case, s, loss)
        if s == 0:
            # gradients of the first step against the fp64 oracle
            P = params_from_golden(g)
            _, G64, _ = O.loss_and_grads(P, *b, masks=masks, drop_p=float(g["dropout"]))
            v = tr.views()
            assert rel_l2(v["user_mlp_grad"].cpu().numpy(), flat_mlp(G64, "user")) <= 1e-5
            gi = v["item_mlp_grad"].cpu().numpy()
            ref = flat_mlp(G64, "item")
            nb2 = G64["item_tower.mlp.3.bias"].size
            assert rel_l2(gi[:-nb2], ref[:-nb2]) <= 1e-5
            assert rel_l2(gi[-nb2:], ref[-nb2:]) <= 1e-4
            for tower in ("user", "item"):
                dense = G64[f"{tower}_tower.embedding.weight"]
                ids = v[f"{tower}_uniq_ids"].cpu().numpy()
                bids = b[0] if tower == "user" else np.concatenate([b[1], b[3]])
                assert np.array_equal(ids, np.unique(bids[bids != 0]))
                assert 0 not in ids                                   # padding row never shows up
                assert np.all(np.diff(ids) > 0)                       # ascending, unique
                assert rel_l2(v[f"{tower}_uniq_grads"].cpu().numpy(), dense[ids]) <= 1e-5
            st = tr.opt_state
            assert abs(st.total_norm - float(g[pre + "total_norm"])) <= 1e-5 * st.total_norm
            assert st.step == 1
        # parameters after the step: every element moves by ~lr (coupled weight decay + Adam normalisation, F5);
        # tolerance 2 % of lr absolute (Adam amplifies 1e-9 gradient noise on near-zero gradients)
        for k, prm in model.state_dict().items():
            ref = g[pre + "after/" + k]
            err = np.abs(prm.detach().cpu().numpy() - ref).max()
            assert err <= 0.25 * lr * (s + 1), (case, s, k, err)
            moved = np.abs(ref - g["init/" + k])
            if k.endswith("embedding.weight"):
                assert (moved > 0).mean() > 0.99                      # dense mode: untouched rows move too
    tr.check_ids()


def test_adam_on_reference_gradients_is_exact(golden):
    """Isolates clip+Adam from gradient rounding: feed the reference's own gradients through the device Adam."""
    import ctypes as C
    import recommendit_b200 as R
    from recommendit_b200 import _lib
    lib = _lib.load()
    g = golden("tt_dup")
    model = model_from_golden(g)
    tr = _trainer(model, g, use_cuda_graph=False)
    ws = _lib.workspace(lib.rb200_sumsq_workspace_bytes(), "cuda")
    moments = {k: (torch.zeros_like(p), torch.zeros_like(p)) for k, p in model.named_parameters()}
    for s in range(2):
        grads = {k: dev(g[f"step{s}/grad/{k}"]) for k, _ in model.named_parameters()}
        _lib.check(lib.rb200_opt_begin_step(tr.opt_dev.data_ptr(), _lib.stream_ptr()))
        items = list(grads.items())
        for i in range(0, len(items), 4):
            segs = (_lib.SumsqSeg * 4)()
            chunk = items[i:i + 4]
            for j, (k, t) in enumerate(chunk):
                segs[j] = _lib.SumsqSeg(t.data_ptr(), t.numel(), None, 0)
            _lib.check(lib.rb200_sumsq_accumulate(tr.opt_dev.data_ptr(), segs, len(chunk), ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
        _lib.check(lib.rb200_grad_norm_clip(tr.opt_dev.data_ptr(), _lib.stream_ptr()))
        st = tr.opt_state
        assert abs(st.total_norm - float(g[f"step{s}/total_norm"])) <= 2e-6 * st.total_norm
        for k, prm in model.named_parameters():
            m, v = moments[k]
            _lib.check(lib.rb200_adam_dense(prm.data_ptr(), grads[k].data_ptr(), m.data_ptr(), v.data_ptr(), prm.numel(),
                                            tr.opt_dev.data_ptr(), _lib.stream_ptr()))
            err = (prm.detach().cpu() - torch.from_numpy(g[f"step{s}/after/{k}"])).abs().max().item()
            assert err <= 2e-3 * float(g["lr"]) + 2e-7, (s, k, err)
            prm.data.copy_(dev(g[f"step{s}/after/{k}"]))


def test_rows_mode_updates_only_touched_rows(golden):
    g = golden("tt_dup")
    model = model_from_golden(g).train()
    before = {k: v.detach().clone() for k, v in model.state_dict().items()}
    tr = _trainer(model, g, adam_mode="rows", use_cuda_graph=False)
    b = batch_from_golden(g, 0)
    tr.load_packed(tr.pack_host(*b))
    tr.step()
    after = model.state_dict()
    for tower, ids in (("user", b[0]), ("item", np.concatenate([b[1], b[3]]))):
        k = f"{tower}_tower.embedding.weight"
        changed = (after[k] != before[k]).any(dim=1).cpu().numpy()
        touched = np.zeros_like(changed); touched[np.unique(ids[ids != 0])] = True
        assert np.array_equal(changed, touched)
        # the touched rows get exactly the dense-mode update
    dense_model = model_from_golden(g).train()
    tr2 = _trainer(dense_model, g, adam_mode="dense", use_cuda_graph=False)
    tr2.load_packed(tr2.pack_host(*b)); tr2.step()
    for tower, ids in (("user", b[0]), ("item", np.concatenate([b[1], b[3]]))):
        k = f"{tower}_tower.embedding.weight"
        rows = dev(np.unique(ids[ids != 0]))
        assert torch.equal(after[k][rows], dense_model.state_dict()[k][rows])
    for k in after:
        if "mlp" in k:
            assert torch.equal(after[k], dense_model.state_dict()[k])


def test_cuda_graph_replay_equals_eager_and_is_deterministic(golden):
    g = golden("tt_dup")
    runs = []
    for use_graph in (False, True, True):
        model = model_from_golden(g).train()
        tr = _trainer(model, g, use_cuda_graph=use_graph)
        losses = []
        for it in range(6):
            b = batch_from_golden(g, it % 2)
            losses.append(tr.step_host(*b))
        runs.append((losses, {k: v.detach().clone() for k, v in model.state_dict().items()}))
        if use_graph:
            assert tr._graph is not None
    for losses, sd in runs[1:]:
        assert losses == runs[0][0]
        for k in sd:
            assert torch.equal(sd[k], runs[0][1][k]), k


@pytest.mark.parametrize("inbatch_mode,graph", [(0, False), (2, False), (None, True)])
def test_in_batch_step_matches_oracle(golden, inbatch_mode, graph):
    """in_batch_bpr_loss step: SIMT kernel, tcgen05 3xTF32 kernel, and the default (auto → tcgen05) under a CUDA graph."""
    g = golden("tt_dup")
    model = model_from_golden(g).train()
    tr = _trainer(model, g, loss="in_batch", inbatch_mode=inbatch_mode, use_cuda_graph=graph)
    b = batch_from_golden(g, 0)
    tr.load_packed(tr.pack_host(*b))
    loss = tr.step().item()
    P = params_from_golden(g)
    l64, G64, _ = O.loss_and_grads(P, *b, in_batch=True)
    assert abs(loss - float(l64)) <= 1e-6
    v = tr.views()
    assert rel_l2(v["user_mlp_grad"].cpu().numpy(), flat_mlp(G64, "user")) <= 1e-5
    P32 = params_from_golden(g, dtype=np.float32)
    S = O.AdamState()
    O.train_step(P32, S, b, lr=float(g["lr"]), in_batch=True)
    for k, prm in model.state_dict().items():
        assert np.abs(prm.detach().cpu().numpy() - P32[k]).max() <= 0.25 * float(g["lr"]), k


def test_item_extra_table_mode_equals_per_sample_genres(golden):
    import recommendit_b200 as R
    g = golden("tt_dup")
    u, p, pg, n, ng = batch_from_golden(g, 0)
    ni = int(g["meta"][1])
    rng = np.random.default_rng(3)
    table = (rng.random((ni + 1, 18)) < 0.2).astype(np.float32)
    m1, m2 = model_from_golden(g).train(), model_from_golden(g).train()
    t1 = _trainer(m1, g, use_cuda_graph=False)
    t2 = _trainer(m2, g, use_cuda_graph=False, item_extra_table=torch.from_numpy(table))
    l1 = t1.step_host(u, p, table[p], n, table[n])
    l2 = t2.step_host(u, p, None, n, None)
    assert l1 == l2
    for k in m1.state_dict():
        assert torch.equal(m1.state_dict()[k], m2.state_dict()[k])
    assert t2.batch_bytes == 3 * len(u) * 8


def test_full_size_step_properties():
    """BASELINE config C2 (B = 8192, ML-1M-shape tables, D = 64, H = 128): size-independent properties."""
    import recommendit_b200 as R
    torch.manual_seed(0)
    nu, ni, B = 6040, 3952, 8192
    model = R.TwoTowerModel(nu, ni, 64, 128, dropout=0.0).cuda().train()
    rng = np.random.default_rng(1)
    u, p, n = rng.integers(1, nu + 1, B), rng.integers(1, ni + 1, B), rng.integers(1, ni + 1, B)
    table = (rng.random((ni + 1, 18)) < 0.1).astype(np.float32)
    P = {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}
    tr = R.FusedBPRTrainer(model, use_cuda_graph=True)
    losses = [tr.step_host(u, p, table[p], n, table[n]) for _ in range(4)]
    assert abs(losses[0] - np.log(2)) < 0.02                      # cosine scores of random towers ≈ 0
    assert losses[3] < losses[0]                                  # same batch repeatedly ⇒ loss goes down
    S = O.AdamState()
    ref = [float(O.train_step(P, S, (u, p, table[p], n, table[n]))[0]) for _ in range(4)]
    assert np.allclose(losses, ref, atol=5e-5), (losses, ref)
    v = tr.views()
    assert len(v["user_uniq_ids"]) == len(np.unique(u)) and len(v["item_uniq_ids"]) == len(np.unique(np.concatenate([p, n])))
    st = tr.opt_state
    assert st.step == 4 and 0 < st.clip_coef <= 1.0
    tr.check_ids()


def test_data_parallel_world1_equals_fused_step(golden):
    """DataParallelBPRTrainer with a single replica is the fused step computed through dense gradients: bit-identical."""
    import recommendit_b200 as R
    g = golden("tt_dup")
    out = []
    for cls in (R.FusedBPRTrainer, R.DataParallelBPRTrainer):
        model = model_from_golden(g).train()
        tr = cls(model, lr=float(g["lr"]), use_cuda_graph=True, seed=3)
        losses = [tr.step_host(*batch_from_golden(g, s % 2)) for s in range(5)]
        out.append((losses, {k: v.detach().clone() for k, v in model.state_dict().items()}))
    assert out[0][0] == out[1][0]
    for k in out[0][1]:
        assert torch.equal(out[0][1][k], out[1][1][k]), k


@pytest.mark.parametrize("B,n_rows,D", [(16385, 5000, 64), (100000, 300, 32), (70000, 2_000_000, 128)])
def test_scatter_rows_large_batches_vs_oracle(B, n_rows, D):
    """Device-wide sort path (> 16384 sample-rows): compact and dense outputs against the oracle, padding id skipped."""
    import ctypes as C
    from recommendit_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(B)
    ids = rng.integers(0, n_rows, B)
    ids[::97] = 0                                                          # padding rows
    rows = rng.standard_normal((B, D)).astype(np.float32)
    d_ids, d_rows = dev(ids), dev(rows)
    uq = torch.empty(B, dtype=torch.int64, device="cuda"); ug = torch.empty(B, D, device="cuda")
    nu = torch.zeros(1, dtype=torch.int32, device="cuda")
    dense = torch.zeros(n_rows, D, device="cuda") if n_rows <= 5000 else None
    slot = torch.full((n_rows,), -1, dtype=torch.int32, device="cuda")
    wsb = lib.rb200_scatter_workspace_bytes(B, n_rows)
    ws = _lib.workspace(wsb, "cuda")
    for _ in range(2):                                                     # twice: deterministic
        if dense is not None:
            dense.zero_()
        _lib.check(lib.rb200_scatter_rows(d_ids.data_ptr(), d_rows.data_ptr(), B, D, n_rows, 0, _lib.ptr(dense), uq.data_ptr(), ug.data_ptr(),
                                          nu.data_ptr(), slot.data_ptr(), ws.data_ptr(), wsb, _lib.stream_ptr()))
        n = int(nu.item())
        got_ids, got = uq[:n].cpu().numpy(), ug[:n].cpu().numpy()
        if _ == 0:
            first = got.copy()
        else:
            assert np.array_equal(first, got)
    exp_ids = np.unique(ids[ids != 0])
    assert np.array_equal(got_ids, exp_ids)
    ref = np.zeros((n_rows, D), np.float64)
    np.add.at(ref, ids[ids != 0], rows[ids != 0].astype(np.float64))
    assert rel_l2(got, ref[exp_ids]) <= 1e-6
    assert np.array_equal(slot.cpu().numpy()[exp_ids], np.arange(n))
    if dense is not None:
        assert rel_l2(dense.cpu().numpy(), ref) <= 1e-6 and torch.count_nonzero(dense[0]) == 0


@pytest.mark.parametrize("B,n_rows,D,kind", [(24576, 1_250_000, 128, "zipf"), (6000, 50_000, 64, "zipf"), (9000, 77, 64, "one"),
                                             (500_000, 4200, 32, "uniform")])
def test_scatter_rows_long_segments(B, n_rows, D, kind):
    """Popular ids (Zipf-skewed users: one id owns hundreds of rows of a batch) go through long_segments_kernel — one CTA per
    segment instead of one warp; a single segment holding the whole batch; more long segments than a block's list holds."""
    from recommendit_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(B)
    if kind == "zipf":
        ids = (rng.zipf(1.05, B) - 1) % n_rows
    elif kind == "one":
        ids = np.full(B, 5, dtype=np.int64)
    else:
        ids = rng.integers(0, n_rows, B)
    rows = rng.standard_normal((B, D)).astype(np.float32)
    d_ids, d_rows = dev(ids), dev(rows)
    uq = torch.empty(B, dtype=torch.int64, device="cuda"); ug = torch.empty(B, D, device="cuda")
    nu = torch.zeros(1, dtype=torch.int32, device="cuda")
    wsb = lib.rb200_scatter_workspace_bytes(B, n_rows)
    ws = _lib.workspace(wsb, "cuda")
    outs = []
    for _ in range(2):
        ug.fill_(float("nan"))
        _lib.check(lib.rb200_scatter_rows(d_ids.data_ptr(), d_rows.data_ptr(), B, D, n_rows, -1, None, uq.data_ptr(), ug.data_ptr(),
                                          nu.data_ptr(), None, ws.data_ptr(), wsb, _lib.stream_ptr()))
        n = int(nu.item())
        outs.append((uq[:n].cpu().numpy().copy(), ug[:n].cpu().numpy().copy()))
    assert np.array_equal(outs[0][1], outs[1][1])                          # deterministic
    exp_ids, counts = np.unique(ids, return_counts=True)
    assert np.array_equal(outs[0][0], exp_ids)
    if kind != "uniform":
        assert counts.max() > 96                                           # the case really has a long segment
    ref = np.zeros((n_rows, D), np.float64)
    np.add.at(ref, ids, rows.astype(np.float64))
    assert rel_l2(outs[0][1], ref[exp_ids]) <= 1e-6


@pytest.mark.parametrize("graph", [False, True])
def test_fused_step_with_popular_ids_matches_fp64_oracle(golden, graph):
    """A batch where one user owns 40 % of the samples and one item 25 % (Zipf-skewed traffic): the long segments are summed by
    a whole block in grad_finish_kernel.  Row gradients, total norm and determinism against the fp64 oracle."""
    g = golden("tt_small")
    nu, ni = int(g["meta"][0]), int(g["meta"][1])
    rng = np.random.default_rng(11)
    B = 700
    u = rng.integers(1, nu + 1, B); u[rng.random(B) < 0.4] = 3
    p = rng.integers(1, ni + 1, B); p[rng.random(B) < 0.25] = 7
    n = rng.integers(1, ni + 1, B); n[rng.random(B) < 0.1] = 7
    pg, ng = (rng.random((B, 18)) < 0.2).astype(np.float32), (rng.random((B, 18)) < 0.2).astype(np.float32)
    b = (u, p, pg, n, ng)
    assert (u == 3).sum() > 200 and ((p == 7).sum() + (n == 7).sum()) > 150
    outs = []
    for rep in range(2):
        model = model_from_golden(g).train()
        model.user_tower.mlp[2].p = model.item_tower.mlp[2].p = 0.0
        tr = _trainer(model, g, use_cuda_graph=graph)
        for _ in range(3 if graph else 1):
            model.load_state_dict({k: torch.from_numpy(np.array(g["init/" + k])) for k in O.PARAM_KEYS})
            tr.load_packed(tr.pack_host(*b))
            tr.step()
        v = tr.views()
        outs.append({k: t.cpu().numpy().copy() for k, t in v.items()})
        if rep == 0:
            P = params_from_golden(g)
            _, G64, _ = O.loss_and_grads(P, *b, masks=None, drop_p=0.0)
            for tower in ("user", "item"):
                dense = G64[f"{tower}_tower.embedding.weight"]
                ids = v[f"{tower}_uniq_ids"].cpu().numpy()
                bids = b[0] if tower == "user" else np.concatenate([b[1], b[3]])
                assert np.array_equal(ids, np.unique(bids[bids != 0]))
                assert rel_l2(v[f"{tower}_uniq_grads"].cpu().numpy(), dense[ids]) <= 1e-5
            if not graph:
                tot = np.sqrt(sum(float((G64[k].astype(np.float64) ** 2).sum()) for k in O.PARAM_KEYS))
                assert abs(tr.opt_state.total_norm - tot) <= 1e-5 * tot
    for k in outs[0]:
        assert np.array_equal(outs[0][k], outs[1][k]), k                  # bit-reproducible


def test_scatter_rows_small_batch_dense_output_with_popular_ids():
    """Single-CTA sort path (<= 16384 rows) writing the DENSE gradient (what the drop-in autograd path asks for) with ids that
    own far more than 8 rows: first row by segment_sum2_kernel, the rest added by long_segments_kernel; padding id skipped."""
    from recommendit_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(2)
    B, n_rows, D = 3000, 400, 64
    ids = rng.integers(0, n_rows, B)
    ids[rng.random(B) < 0.3] = 7
    ids[rng.random(B) < 0.1] = 0                                           # padding rows
    rows = rng.standard_normal((B, D)).astype(np.float32)
    d_ids, d_rows = dev(ids), dev(rows)
    uq = torch.empty(B, dtype=torch.int64, device="cuda"); ug = torch.empty(B, D, device="cuda")
    nu = torch.zeros(1, dtype=torch.int32, device="cuda")
    dense = torch.zeros(n_rows, D, device="cuda")
    wsb = lib.rb200_scatter_workspace_bytes(B, n_rows)
    ws = _lib.workspace(wsb, "cuda")
    _lib.check(lib.rb200_scatter_rows(d_ids.data_ptr(), d_rows.data_ptr(), B, D, n_rows, 0, dense.data_ptr(), uq.data_ptr(), ug.data_ptr(),
                                      nu.data_ptr(), None, ws.data_ptr(), wsb, _lib.stream_ptr()))
    n = int(nu.item())
    ref = np.zeros((n_rows, D), np.float64)
    np.add.at(ref, ids[ids != 0], rows[ids != 0].astype(np.float64))
    exp_ids = np.unique(ids[ids != 0])
    assert (ids == 7).sum() > 500
    assert np.array_equal(uq[:n].cpu().numpy(), exp_ids)
    assert rel_l2(ug[:n].cpu().numpy(), ref[exp_ids]) <= 1e-6
    assert rel_l2(dense.cpu().numpy(), ref) <= 1e-6 and torch.count_nonzero(dense[0]) == 0


def test_train_epoch_pipelined_equals_step_host_loop(golden):
    """train_epoch (H2D prefetch on a copy stream, async loss read-back) must produce exactly the parameters and the mean
    loss of the synchronous per-step loop (train_embeddings.py:170-199) on the same batches."""
    g = golden("tt_dup")
    rng = np.random.default_rng(0)
    nu, ni, B = int(g["meta"][0]), int(g["meta"][1]), 256

    def batch():
        return (rng.integers(0, nu + 1, B), rng.integers(0, ni + 1, B), (rng.random((B, 18)) < 0.2).astype(np.float32),
                rng.integers(0, ni + 1, B), (rng.random((B, 18)) < 0.2).astype(np.float32))
    batches = [batch() for _ in range(7)]
    m1, m2 = model_from_golden(g).train(), model_from_golden(g).train()
    t1, t2 = _trainer(m1, g, seed=5), _trainer(m2, g, seed=5)
    losses = [t1.step_host(*b) for b in batches]
    packed = [t2.pack_host(*b).clone().pin_memory() for b in batches]
    mean = t2.train_epoch(packed)
    assert abs(mean - float(np.mean(losses))) <= 1e-7
    for (k, a), (_, b) in zip(m1.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    assert np.isnan(t2.train_epoch([]))


@pytest.mark.gpu
@pytest.mark.parametrize("B,n_rows,D,kind", [(37000, 700000, 128, "zipf"), (3000, 500, 64, "uniform"), (20000, 90000, 128, "empty_slots")])
def test_scatter_plan_then_apply_equals_scatter_rows(B, n_rows, D, kind):
    """Two-phase segment sum (rb200_scatter_plan on the ids — on another stream, before the rows exist — then rb200_scatter_apply):
    bit-identical to rb200_scatter_rows, including empty slots (id -1: the padded exchange's unused bucket slots) and popular ids."""
    from recommendit_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(B)
    ids = (rng.zipf(1.05, B) - 1) % n_rows if kind == "zipf" else rng.integers(0, n_rows, B)
    if kind == "empty_slots":
        ids[rng.random(B) < 0.5] = -1
    d_ids = dev(ids)
    wsb = lib.rb200_scatter_workspace_bytes(B, n_rows)
    uq1 = torch.empty(B, dtype=torch.int64, device="cuda"); ug1 = torch.full((B, D), float("nan"), device="cuda")
    nu1 = torch.zeros(1, dtype=torch.int32, device="cuda")
    uq2 = torch.empty_like(uq1); ug2 = torch.full((B, D), float("nan"), device="cuda"); nu2 = torch.zeros_like(nu1)
    ws1, ws2 = _lib.workspace(wsb, "cuda"), _lib.workspace(wsb, "cuda")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                                          # the plan: ids only, side stream
        _lib.check(lib.rb200_scatter_plan(d_ids.data_ptr(), B, n_rows, -1, uq2.data_ptr(), nu2.data_ptr(), None, ws2.data_ptr(), wsb,
                                          _lib.stream_ptr()))
    rows = rng.standard_normal((B, D)).astype(np.float32)                  # the rows come later
    d_rows = dev(rows)
    _lib.check(lib.rb200_scatter_rows(d_ids.data_ptr(), d_rows.data_ptr(), B, D, n_rows, -1, None, uq1.data_ptr(), ug1.data_ptr(),
                                      nu1.data_ptr(), None, ws1.data_ptr(), wsb, _lib.stream_ptr()))
    torch.cuda.current_stream().wait_stream(side)
    _lib.check(lib.rb200_scatter_apply(d_rows.data_ptr(), B, D, n_rows, None, uq2.data_ptr(), ug2.data_ptr(), nu2.data_ptr(), ws2.data_ptr(),
                                       wsb, _lib.stream_ptr()))
    n1, n2 = int(nu1.item()), int(nu2.item())
    assert n1 == n2 == len(np.unique(ids[ids >= 0]))
    assert torch.equal(uq1[:n1], uq2[:n2]) and torch.equal(ug1[:n1], ug2[:n2])
    exp = np.zeros((n_rows, D), np.float64)
    np.add.at(exp, ids[ids >= 0], rows[ids >= 0].astype(np.float64))
    got = ug2[:n2].cpu().numpy()
    np.testing.assert_allclose(got, exp[uq2[:n2].cpu().numpy()], rtol=0, atol=5e-4 if kind == "zipf" else 1e-5)
