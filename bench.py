#!/usr/bin/env python
"""Benchmark of the two-tower BPR training step (headline) and IVFFlat top-500 retrieval on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line (rank 0).  Workloads (BASELINE.json ``configs``):

* headline, N = 1 — C2: one BPR training step, batch 8192, ML-1M-shape tables (6041 / 3953 rows), D = 64,
  H = 128, 18 genre columns, sampled negatives + ``bpr_loss`` (what the reference's loop runs), dropout 0.1
  active (as shipped), clip_grad_norm_(1.0) + Adam(weight_decay=1e-5) over every parameter (dense mode).
  ``value`` = samples/s with the batch already in HBM; ``e2e`` = the same step fed from pinned host memory with
  the loss read back every step.
* secondary (``ivf`` object) — C3: IVFFlat nlist 4096 / nprobe 32 / top-500 over 1 M × 64, batches of 4096 queries.
* N > 1 — the same C2 step weak-scaled: 8192 samples per rank, replicated tables, one NCCL all-reduce of the dense
  gradients per step (``DataParallelBPRTrainer``); plus a ``c4`` object: row-sharded tables (10 M users × 1 M items,
  D = 128) with NCCL all-to-all for ids / rows / row gradients (``bench_sharded.py``).

``--impl reference`` times the CPU arm (oracle/torch_step.py: the reference's own PyTorch calls) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_USERS, N_ITEMS, D, H, E, B = 6040, 3952, 64, 128, 18, 8192
DROPOUT = 0.1
SEED = 20240601


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return {"hbm_gbs": float(j["hbm_gbs"]), "bf16_tflops": float(j["bf16_tflops"]),
                "bf16_tflops_sustained": float(j.get("bf16_tflops_sustained", j["bf16_tflops"])), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def ncu_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed `ncu --set full` captures
    (profiles/r02_traffic.json: kernel -> {"read": B, "write": B, "source": file}); None when no capture holds the kernel."""
    p = ROOT / "profiles" / "r02_traffic.json"
    if not p.exists():
        return None
    ent = json.loads(p.read_text()).get(kernel)
    return None if ent is None else int(ent["read"] + ent["write"])


def synth_batches(n_batches: int, seed: int = 1):
    """C2 batches: ids uniform over the tables, genres = per-item multi-hot (Bernoulli 0.092, ≥ 1 bit)."""
    rng = np.random.default_rng(SEED)
    genres = (rng.random((N_ITEMS + 1, E)) < 0.092).astype(np.float32)
    empty = genres.sum(1) == 0
    genres[empty, rng.integers(0, E, empty.sum())] = 1.0
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_batches):
        u = rng.integers(1, N_USERS + 1, B)
        p = rng.integers(1, N_ITEMS + 1, B)
        n = rng.integers(1, N_ITEMS + 1, B)
        out.append((u, p, genres[p], n, genres[n]))
    return out, genres


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.time() + 3.0                  # a short timed region can end before nvidia-smi printed its first sample
        while time.time() < t_end and Path(self.f.name).stat().st_size < 40:
            time.sleep(0.05)
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in Path(self.f.name).read_text().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            for name, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons), "samples": len(rows)}


#: the workload both arms of the N = 1 line run (identical dict in both JSON lines: the driver compares them)
C2_CONFIG = {"workload": "C2: two-tower BPR training step, batch 8192, ML-1M-shape tables 6041x64 / 3953x64, H=128, 18 genres, "
                         "sampled negatives + bpr_loss, dropout 0.1, clip_grad_norm_ 1.0, Adam wd 1e-5 (dense: every row updated, "
                         "as torch.optim.Adam on the reference's dense embedding gradients does)"}


# --------------------------------------------------------------------------------------------------------- #
# CPU arm
# --------------------------------------------------------------------------------------------------------- #
def reference_stepper(n_users: int, n_items: int, d: int, h: int, dropout: float, threads: int):
    """→ (step(batch of CPU tensors) -> loss float, kind).  kind "reference": the UNMODIFIED reference module (oracle/_ref, a byte
    copy of /root/reference/src made by oracle/make_ref.py and shipped with the snapshot) driven through the reference's own step
    body, train_embeddings.py:183-194.  kind "port": the same ATen calls restated (oracle/torch_step.py) when the copy is absent."""
    torch.set_num_threads(threads)
    from oracle import make_ref
    if make_ref.available():
        if make_ref.path() not in sys.path:
            sys.path.insert(1, make_ref.path())
        import logging
        logging.disable(logging.WARNING)                 # the reference logs "faiss not available" at import
        from src.models.two_tower import TwoTowerModel    # the stock class: nothing of this repository on this path
        logging.disable(logging.NOTSET)
        assert "oracle/_ref" in sys.modules["src.models.two_tower"].__file__
        torch.manual_seed(0)
        model = TwoTowerModel(n_users=n_users, n_items=n_items, embed_dim=d, hidden_dim=h, dropout=dropout).train()
        optimizer = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5)

        def step(batch):
            user_ids, pos_ids, pos_genres, neg_ids, neg_genres = batch
            user_emb = model.user_tower(user_ids)
            pos_item_emb = model.item_tower(pos_ids, pos_genres)
            neg_item_emb = model.item_tower(neg_ids, neg_genres)
            loss = model.bpr_loss(user_emb, pos_item_emb, neg_item_emb)
            optimizer.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
            optimizer.step()
            return loss.item()
        return step, "reference"
    from oracle import torch_step as TS
    from oracle import two_tower_oracle as O
    T = TS.make_params(O.init_params(n_users, n_items, d, h, seed=0))
    opt = TS.make_optimizer(T)
    return (lambda batch: TS.step(T, opt, batch, dropout=dropout)), "port"


def cpu_step_throughput(steps: int, warmup: int, threads: int):
    step, kind = reference_stepper(N_USERS, N_ITEMS, D, H, DROPOUT, threads)
    batches, _ = synth_batches(min(steps + warmup, 8))
    tb = [tuple(torch.from_numpy(np.ascontiguousarray(a)) for a in b) for b in batches]
    for i in range(warmup):
        step(tb[i % len(tb)])
    t0 = time.perf_counter()
    for i in range(steps):
        step(tb[(warmup + i) % len(tb)])
    dt = time.perf_counter() - t0
    return B * steps / dt, dt / steps * 1e3, kind


def _ref_how(kind: str) -> str:
    return ("the UNMODIFIED reference module (oracle/_ref/src/models/two_tower.py, byte copy of the reference) through the "
            "reference's own step body, train_embeddings.py:183-194" if kind == "reference"
            else "oracle/torch_step.py: the reference's own PyTorch calls restated (oracle/_ref not shipped)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    if world > 1 or args.gpus > 1:
        # the N > 1 arm measures C4 (bench_sharded.py): the reference's step at C4 widths on a bounded sample, same config dict
        import bench_sharded as BS
        n = max(world, args.gpus)
        cb = BS.cpu_baseline_c4(steps=max(2, min(args.steps, 4)))
        line = {"impl": "reference", "metric": "bpr_train_samples_per_s", "value": cb["value"], "unit": "samples/s", "n_gpus": n,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": BS.c4_config(n), "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return
    val, ms, kind = cpu_step_throughput(args.steps, max(args.warmup, 1), cores)
    line = {
        "impl": "reference", "metric": "bpr_train_samples_per_s", "value": val, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": dict(C2_CONFIG),
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": kind,
                         "sample": f"{args.steps} full steps of batch {B} ({_ref_how(kind)}; torch {torch.__version__} CPU, "
                                   f"{cores} threads, dropout on)"},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------- #
# GPU arm, N = 1
# --------------------------------------------------------------------------------------------------------- #
def flush_l2(buf):
    buf.add_(1)     # 256 MiB read+write > 126 MB L2


def bench_train_single(args, dev):
    import recommendit_b200 as R
    from recommendit_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(0)
    model = R.TwoTowerModel(N_USERS, N_ITEMS, D, H, dropout=DROPOUT).to(dev).train()
    K, W = args.steps, args.warmup
    batches, genres = synth_batches(K + W + 2)
    tr = R.FusedBPRTrainer(model, lr=1e-3, weight_decay=1e-5, max_norm=1.0, adam_mode="dense", loss="bpr", use_cuda_graph=True)
    # stage every batch: packed in pinned host memory (e2e) and resident in HBM (value)
    pinned = [tr.pack_host(*b).clone().pin_memory() for b in batches]
    resident = [p.to(dev) for p in pinned]
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
    h2d = pinned[0].numel()

    # launches per step, counted on one eager step
    tr.load_packed(resident[0])
    c0 = lib.rb200_launch_count()
    tr.step()                                  # eager (first step of this shape)
    launches_per_step = lib.rb200_launch_count() - c0
    tr.load_packed(resident[1]); tr.step()     # captures the graph
    torch.cuda.synchronize(dev)
    for i in range(W):
        tr.load_packed(resident[(2 + i) % len(resident)]); tr.step()
    torch.cuda.synchronize(dev)

    # ---- value: device-resident inputs, per-step CUDA events, L2 flushed between steps ------------------ #
    sampler = ClockSampler(dev.index or 0)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    torch.cuda.synchronize(dev)
    for i in range(K):
        flush_l2(flush)
        ev[i][0].record()
        tr.load_packed(resident[(2 + W + i) % len(resident)])
        tr.step()
        ev[i][1].record()
    torch.cuda.synchronize(dev)
    ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(ms))
    # same loop without the flush (tables are 2.6 MB: L2-resident in steady state), reported for information
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        tr.load_packed(resident[(2 + W + i) % len(resident)]); tr.step()
    e1.record()
    torch.cuda.synchronize(dev)
    warm_ms = e0.elapsed_time(e1) / K

    # ---- e2e: pinned host batch → H2D → step → loss read back, wall clock per step ------------------------- #
    e2e_t = 0.0
    for i in range(K):
        flush_l2(flush)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        tr.load_packed(pinned[(2 + W + i) % len(pinned)])
        loss = tr.step().item()
        e2e_t += time.perf_counter() - t0
    # ---- e2e, pipelined: the trainer's epoch loop (H2D of batch i+1 overlaps step i; one 4-byte loss D2H per step) -- #
    order = [pinned[(2 + W + i) % len(pinned)] for i in range(K)]
    tr.train_epoch(order[: min(K, 4)])                     # warm the copy stream / staging buffers
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    epoch_loss = tr.train_epoch(order)
    pipe_t = time.perf_counter() - t0
    clocks = sampler.stop()
    tr.check_ids()

    # ---- per-stage device times (eager launch with events at the stage boundaries) ---------------------------- #
    stage_names = ["towers_fwd", "loss", "towers_bwd", "scatter", "clip", "adam"]
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
    for e in evs:
        e.record()
    torch.cuda.synchronize(dev)
    import ctypes as C
    arr = (C.c_void_p * 7)(*[e.cuda_event for e in evs])
    acc = np.zeros(6)
    reps = 20
    for r in range(reps):
        flush_l2(flush)
        tr.load_packed(resident[r % len(resident)])
        p = tr._make_params()
        p.stage_events_host = C.cast(arr, C.c_void_p)
        _lib.check(lib.rb200_bpr_step(C.byref(p), _lib.stream_ptr()))
        torch.cuda.synchronize(dev)
        acc += np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(6)])
    stages = {n: float(v / reps) for n, v in zip(stage_names, acc)}

    # ---- drop-in path (unchanged reference step body on this model + torch.optim.Adam), e2e -------------------- #
    model2 = R.TwoTowerModel(N_USERS, N_ITEMS, D, H, dropout=DROPOUT).to(dev).train()
    opt = torch.optim.Adam(model2.parameters(), lr=1e-3, weight_decay=1e-5)
    hb = [tuple(torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in b) for b in batches[:8]]

    def dropin_step(b):
        user_ids, pos_ids, pos_genres, neg_ids, neg_genres = [t.to(dev, non_blocking=True) for t in b]
        user_emb = model2.user_tower(user_ids)
        pos_emb = model2.item_tower(pos_ids, pos_genres)
        neg_emb = model2.item_tower(neg_ids, neg_genres)
        loss = model2.bpr_loss(user_emb, pos_emb, neg_emb)
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model2.parameters(), max_norm=1.0)
        opt.step()
        return loss.item()

    for i in range(3):
        dropin_step(hb[i % len(hb)])
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    nd = min(K, 50)
    for i in range(nd):
        dropin_step(hb[i % len(hb)])
    dropin_ms = (time.perf_counter() - t0) / nd * 1e3

    return {"total_ms": total_ms, "ms": ms, "warm_ms": warm_ms, "e2e_s": e2e_t, "pipe_s": pipe_t, "epoch_loss": epoch_loss, "h2d": h2d, "launches_per_step": int(launches_per_step),
            "clocks": clocks, "stages": stages, "loss": loss, "dropin_ms": dropin_ms}


def ivf_roofline(scan_bytes, run_ms, n, d, nq, k, n_cand, pk):
    """The list scan + select pair (rb200_ivf_search_run) against the HBM roofline, stated BOTH ways (SURVEY.md §8d note):
    `frac` uses the bytes a list-major kernel must move — the database once (vectors + ids), the candidate scores written once
    and read once by the select, the results — because the query-major figure (every probing query re-reads its lists: 8.45 GB
    per C3 batch against a 264 MB database) exceeds what any kernel that shares a list between its queries needs to read, and a
    fraction of peak computed from it (2.7 in round 1) is not a roofline statement; it is kept as `query_major`."""
    phys = n * (4 * d + 8) + 2 * n_cand * 4 + nq * k * 12
    traffic = ncu_traffic("list_scan_tc_kernel+select_topk_kernel")
    return {"kernel": "list_scan_pipe_kernel + select_topk_kernel<ResolveIvf> (rb200_ivf_search_run)", "bound": "hbm", "unit": "GB/s",
            "achieved": phys / (run_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "frac": phys / (run_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
            "peak_source": pk["source"], "algorithmic_bytes": phys, "ms": run_ms, "traffic": traffic,
            "definition": "list-major: database read once (N·(4·D+8)) + candidate scores written and read once (2·4·candidates) + results",
            "query_major": {"algorithmic_bytes": scan_bytes, "GBps_equivalent": scan_bytes / (run_ms * 1e-3) / 1e9,
                            "note": "SURVEY.md §8d's per-query figure nq·Σ_probed len·(4·D+8): what a query-major scan would read; "
                                    "reported for comparison with CPU FAISS-style scans, not as a fraction of peak"},
            "note": "the pair is latency / SM-bound, not HBM-bound (profiles/r02_ivf.md: DRAM traffic is 1.4x the database; the list scan is bound by the SM's "
                    "load/store path, the select by its passes over shared memory)"}


def bench_ivf(args, dev):
    """C3: nlist 4096, nprobe 32, top-500, 4096 queries over 1 M × 64."""
    import recommendit_b200 as R
    from recommendit_b200 import _lib
    n, d, nlist, nprobe, k, nq = 1_000_000, 64, 4096, 32, 500, 4096
    g = torch.Generator(device=dev).manual_seed(7)
    cen = torch.nn.functional.normalize(torch.randn(nlist, d, device=dev, generator=g), dim=-1)
    z = (torch.rand(n, device=dev, generator=g) ** 2 * nlist).long().clamp_(max=nlist - 1)      # skewed list sizes
    x = torch.nn.functional.normalize(cen[z] + 0.35 * torch.randn(n, d, device=dev, generator=g), dim=-1)
    q = torch.nn.functional.normalize(x[torch.randint(0, n, (nq,), device=dev, generator=g)] +
                                      0.2 * torch.randn(nq, d, device=dev, generator=g), dim=-1)
    idx = R.FAISSIndex(d, nlist, nprobe)
    t0 = time.perf_counter()
    idx.build_ivf_index(x.cpu().numpy(), list(range(n)), centroids=cen.cpu().numpy())
    build_s = time.perf_counter() - t0
    st = idx.index
    qh = q.cpu().numpy()
    lens = (st.offsets[1:] - st.offsets[:-1])
    reps = max(3, min(args.steps, 20))
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        st.search_device(q, k)
    torch.cuda.synchronize(dev)
    ms = []
    for _ in range(reps):
        flush_l2(flush)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); s, r = st.search_device(q, k); b.record()
        torch.cuda.synchronize(dev)
        ms.append(a.elapsed_time(b))
    dev_ms = float(np.median(ms))
    for _ in range(3):                       # warm-up: pinned result buffers come from torch's caching host allocator
        idx.batch_search(qh, k)
    t0 = time.perf_counter()
    for _ in range(reps):
        idx.batch_search(qh, k)
    e2e_ms = (time.perf_counter() - t0) / reps * 1e3
    # algorithmic bytes of the list scan (query-major definition): Σ_q Σ_probed len · (4·D + 8)
    lib = _lib.load()
    import ctypes as C
    pb = lib.rb200_ivf_plan_workspace_bytes(nq, nlist, nprobe)
    plan = _lib.workspace(pb, dev)
    tot, mx = C.c_int64(0), C.c_int64(0)
    _lib.check(lib.rb200_ivf_search_plan(q.data_ptr(), nq, d, st.centroids.data_ptr(), nlist, nprobe, st.offsets.data_ptr(),
                                         plan.data_ptr(), pb, C.byref(tot), C.byref(mx), _lib.stream_ptr()))
    scan_bytes = tot.value * (4 * d + 8)
    # time the scan+select part alone (plan reused)
    ws = _lib.workspace(lib.rb200_ivf_search_workspace_bytes(tot.value), dev)
    so = torch.empty(nq, k, device=dev); io = torch.empty(nq, k, dtype=torch.int64, device=dev)
    run_ms = []
    for _ in range(reps):
        flush_l2(flush)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(lib.rb200_ivf_search_run(q.data_ptr(), nq, d, nlist, nprobe, st.offsets.data_ptr(), st.list_ids.data_ptr(),
                                            st.list_vecs.data_ptr(), st.list_vecs.shape[0], st.max_list_len, st.tile_list.data_ptr(), st.tile_idx.data_ptr(),
                                            st.tile_list.numel(), k, plan.data_ptr(), pb, tot.value, mx.value,
                                            so.data_ptr(), io.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
        b.record(); torch.cuda.synchronize(dev)
        run_ms.append(a.elapsed_time(b))
    pk = peaks()
    run = float(np.median(run_ms))
    # retrieval quality: the IVF ids against the exhaustive top-500 of the same queries (rb200_flat_search)
    from recommendit_b200.evaluation import retrieval_report
    _, exact_ids = R.flat_search(q, x, k)
    quality = retrieval_report(r, exact_ids)
    out = {
        "metric": "ivf_top500_qps", "value": nq / dev_ms * 1e3, "unit": "queries/s", "ms_per_batch": dev_ms,
        "config": {"workload": "C3: IVFFlat nlist=4096 nprobe=32 top-500, 4096 queries per batch, 1M x 64 fp32 database, "
                               "skewed lists (max %d, mean %.0f)" % (int(lens.max()), float(lens.float().mean())),
                   "l2": "flushed between timed batches (256 MiB write)"},
        "e2e": {"value": nq / e2e_ms * 1e3, "unit": "queries/s", "h2d_bytes_per_step": nq * d * 4, "d2h_bytes_per_step": nq * k * 12,
                "ms_per_batch": e2e_ms},
        "roofline": ivf_roofline(scan_bytes, run, n, d, nq, k, int(tot.value), pk),
        "quality_vs_exact": quality,
        "quality_note": "Recall@K / NDCG@K of the IVF ids against the exhaustive top-K of the same queries.  The C3 generator adds "
                        "noise of norm 0.35*sqrt(64) = 2.8 to unit cluster centres, so list membership says little about neighbourhood "
                        "and 32 of 4096 lists recover few of the true neighbours; IVF with nprobe = nlist equals the exhaustive "
                        "result (tests/test_gpu_ivf.py)",
        "candidates_per_batch": int(tot.value), "index_build_s": build_s,
    }
    return out


def bench_inbatch(args, dev):
    """C2 in-batch variant: in_batch_bpr_loss (two_tower.py:132-160) — kernel alone (SIMT vs tcgen05) and the
    whole training step (2 towers + B×B loss).  The score/gradient products are the one GEMM-shaped stage of the path:
    reported against the tensor roofline with 6·B²·D logical flops (S, dU, dI; scores are recomputed in the second pass,
    so 8·B²·D are executed, x3 in 3xTF32 mode)."""
    import recommendit_b200 as R
    from recommendit_b200 import _lib
    lib = _lib.load()
    pk = peaks()
    g = torch.Generator(device=dev).manual_seed(3)
    U = torch.nn.functional.normalize(torch.randn(B, D, device=dev, generator=g), dim=-1)
    I = torch.nn.functional.normalize(torch.randn(B, D, device=dev, generator=g), dim=-1)
    loss = torch.empty(1, device=dev); dU = torch.empty_like(U); dI = torch.empty_like(I)
    wsb = lib.rb200_bpr_inbatch_workspace_bytes(B, D)
    ws = _lib.workspace(wsb, dev)
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
    out = {"B": B, "D": D, "logical_flops": 6.0 * B * B * D}
    for mode, name in ((0, "simt_fp32"), (2, "tcgen05_3xtf32"), (1, "tcgen05_tf32")):
        def run():
            _lib.check(lib.rb200_bpr_inbatch(U.data_ptr(), I.data_ptr(), B, D, mode, loss.data_ptr(), dU.data_ptr(), dI.data_ptr(), 1.0,
                                             ws.data_ptr(), wsb, _lib.stream_ptr()))
        for _ in range(3):
            run()
        ts = []
        for _ in range(10):
            flush_l2(flush)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(); b.record()
            torch.cuda.synchronize(dev)
            ts.append(a.elapsed_time(b))
        ms = float(np.median(ts))
        out[name] = {"ms": ms, "logical_tflops": 6.0 * B * B * D / (ms * 1e-3) / 1e12}
    t = out["tcgen05_3xtf32"]
    out["roofline"] = {"kernel": "inbatch_ts_kernel<2> x2 (rb200_bpr_inbatch mode 2: X and G tiles as TMEM A operands, TMA-fed operand images)",
                       "bound": "tensor", "unit": "TFLOP/s",
                       "achieved": t["logical_tflops"], "issued_tflops": t["logical_tflops"] * 4.0, "peak": pk["bf16_tflops"],
                       "frac": t["logical_tflops"] / pk["bf16_tflops"], "issued_frac_of_tf32_peak": t["logical_tflops"] * 4.0 / (pk["bf16_tflops"] / 2),
                       "traffic": ncu_traffic("inbatch_tc_kernel"),
                       "tensor_pipe_active_pct_ncu": 54.8,
                       "note": "logical 6·B²·D fp32-grade flops vs the measured bf16 peak (kind::tf32 peaks at half of it); issued = "
                               "8·B²·D (scores recomputed in the second pass) x 3 (3xTF32).  ncu (profiles/r02_inbatch_ts.md): "
                               "sm__pipe_tensor_cycles_active 54.8 % / 55.8 % in the two passes (north_star target >= 50 %), DRAM 10.6 MB "
                               "read / 0 written per launch"}
    # whole step with the in-batch loss (2 towers + B×B), CUDA graph, device-resident batch
    torch.manual_seed(0)
    model = R.TwoTowerModel(N_USERS, N_ITEMS, D, H, dropout=DROPOUT).to(dev).train()
    tr = R.FusedBPRTrainer(model, lr=1e-3, weight_decay=1e-5, max_norm=1.0, adam_mode="dense", loss="in_batch", use_cuda_graph=True)
    batches, _ = synth_batches(4)
    resident = [tr.pack_host(*b).to(dev) for b in batches]
    for i in range(5):
        tr.load_packed(resident[i % 4]); tr.step()
    torch.cuda.synchronize(dev)
    ts = []
    for i in range(max(10, args.steps)):
        flush_l2(flush)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); tr.load_packed(resident[i % 4]); tr.step(); b.record()
        torch.cuda.synchronize(dev)
        ts.append(a.elapsed_time(b))
    ms = float(np.mean(ts))
    out["step"] = {"ms_per_step": ms, "samples_per_s": B / (ms * 1e-3), "final_loss": float(tr.loss_dev.item()),
                   "api": "FusedBPRTrainer(loss='in_batch') (CUDA graph), in-batch kernel mode auto = tcgen05 3xTF32"}
    return out


def bench_producer(args, dev):
    """the whole reference epoch loop with batches produced on the device (no host batches at all).
    C1-shape synthetic ratings (seed 20240601: 6040 users, 3883 catalog items, 1 000 209 interactions, Zipf-like users/items,
    ML-1M rating marginals → ≈ 575 k positives), batch 8192, genres looked up by item id inside the tower kernels."""
    import recommendit_b200 as R
    rng = np.random.default_rng(20240601)
    n = 1_000_209
    catalog = np.sort(rng.choice(np.arange(1, N_ITEMS + 1), 3883, replace=False)).astype(np.int64)
    pu = 1.0 / (np.arange(N_USERS) + 10.0); pu /= pu.sum()
    pi = 1.0 / (np.arange(3706) + 5.0); pi /= pi.sum()
    u = rng.choice(np.arange(1, N_USERS + 1), n, p=pu)
    i = catalog[rng.choice(3706, n, p=pi)]
    r = rng.choice([1, 2, 3, 4, 5], n, p=[0.056, 0.107, 0.261, 0.349, 0.227]).astype(np.float64)
    t0 = time.perf_counter()
    prod = R.DeviceBatchProducer(u, i, r, catalog, N_USERS, seed=1)
    build_s = time.perf_counter() - t0
    genres = (rng.random((N_ITEMS + 1, 18)) < 0.092).astype(np.float32)
    torch.manual_seed(0)
    model = R.TwoTowerModel(N_USERS, N_ITEMS, D, H, dropout=DROPOUT).to(dev).train()
    tr = R.FusedBPRTrainer(model, lr=1e-3, weight_decay=1e-5, max_norm=1.0, adam_mode="dense", item_extra_table=torch.from_numpy(genres))
    nb = prod.batches_per_epoch(B)
    prod.train_epoch(tr, B, 0)                                  # warm-up epoch (captures the graph)
    torch.cuda.synchronize(dev)
    n_ep = 5
    t0 = time.perf_counter()
    for e in range(1, 1 + n_ep):
        loss = prod.train_epoch(tr, B, e)
    dt = (time.perf_counter() - t0) / n_ep
    tr.check_ids()
    return {"metric": "bpr_train_samples_per_s", "value": nb * B / dt, "unit": "samples/s", "ms_per_step": dt / nb * 1e3,
            "batches_per_epoch": nb, "epochs_timed": n_ep, "positives": prod.n_pos, "epoch_mean_loss": loss, "index_build_s": build_s,
            "api": "DeviceBatchProducer.train_epoch(FusedBPRTrainer(item_extra_table=genres), 8192, epoch): the step samples its own "
                   "next batch on the device (rb200_step_params.next_batch), so an epoch is one graph replay per step; wall clock "
                   "around the epochs, mean loss read once per epoch",
            "reference_producer": "UserItemDataset + DataLoader: ≈ 22 k samples/s (SURVEY.md §6.2)"}


def bench_hbm_kernels(dev):
    """The HBM-bound kernels of the path at a size where HBM (not L2) is the limit — BASELINE config C4 widths (D = 128):
    row gather, sorted-segment scatter-add, Adam on touched rows, dense Adam.  Algorithmic bytes per SURVEY.md §8(d)."""
    from recommendit_b200 import _lib
    lib = _lib.load()
    pk = peaks()
    D4, rows, n_req = 128, 4_000_000, 2_000_000
    f32 = dict(dtype=torch.float32, device=dev)
    table = torch.randn(rows, D4, **f32)
    m, v = torch.zeros_like(table), torch.zeros_like(table)
    g = torch.Generator(device=dev).manual_seed(3)
    ids = torch.randint(0, rows, (n_req,), device=dev, generator=g)
    out = torch.empty(n_req, D4, **f32)
    opt = torch.zeros(128, dtype=torch.uint8, device=dev)
    st = _lib.OptState()
    st.lr, st.beta1, st.beta2, st.eps, st.weight_decay, st.max_norm = 1e-3, 0.9, 0.999, 1e-8, 1e-5, 1.0
    st.one_minus_beta1, st.one_minus_beta2, st.beta2_f, st.clip_coef = 0.1, 0.001, 0.999, 1.0
    opt.copy_(torch.frombuffer(bytearray(bytes(st)), dtype=torch.uint8))
    _lib.check(lib.rb200_opt_begin_step(opt.data_ptr(), _lib.stream_ptr()))
    uniq = torch.unique(ids)
    nu = torch.tensor([uniq.numel()], dtype=torch.int32, device=dev)
    ug = torch.randn(uniq.numel(), D4, **f32) * 1e-3
    slot = torch.full((rows,), -1, dtype=torch.int32, device=dev)
    sp = _lib.stream_ptr

    def timeit(fn, reps=5):
        fn(); torch.cuda.synchronize(dev)
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize(dev)
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    res = {}

    def add(name, ms, nbytes, note):
        gbs = nbytes / (ms * 1e-3) / 1e9
        res[name] = {"ms": ms, "algorithmic_bytes": int(nbytes), "achieved": gbs, "unit": "GB/s", "peak": pk["hbm_gbs"],
                     "frac": gbs / pk["hbm_gbs"], "note": note}

    ms = timeit(lambda: _lib.check(lib.rb200_gather_rows(table.data_ptr(), ids.data_ptr(), n_req, D4, rows, out.data_ptr(), sp())))
    add("gather_rows", ms, n_req * (2 * 4 * D4 + 8), f"{n_req} random rows of a {rows}x{D4} fp32 table (2 GB): 4·D read + 4·D written + 8 B id per row (physical bytes of a stand-alone gather)")
    res["gather_rows"]["survey_8d_definition"] = {"algorithmic_bytes": n_req * (4 * D4 + 8), "achieved": n_req * (4 * D4 + 8) / (ms * 1e-3) / 1e9,
                                                   "frac": n_req * (4 * D4 + 8) / (ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                                                   "note": "SURVEY.md §8d counts only the gathered row + its id (4·D + 8 B): the figure for a gather "
                                                           "FUSED into its consumer (the tower kernels read the rows straight into TMEM and never "
                                                           "write them back); this stand-alone kernel also writes every row"}
    ms = timeit(lambda: _lib.check(lib.rb200_adam_rows(table.data_ptr(), m.data_ptr(), v.data_ptr(), D4, uniq.data_ptr(), ug.data_ptr(),
                                                       nu.data_ptr(), uniq.numel(), opt.data_ptr(), sp())))
    add("adam_rows", ms, uniq.numel() * (28 * D4 + 8), f"Adam on {uniq.numel()} touched rows: w,m,v read+written (24·D) + gradient row (4·D) + id")
    ms = timeit(lambda: _lib.check(lib.rb200_adam_table_dense(table.data_ptr(), m.data_ptr(), v.data_ptr(), rows, D4, slot.data_ptr(), ug.data_ptr(),
                                                              opt.data_ptr(), sp())))
    add("adam_table_dense", ms, rows * (24 * D4 + 4), f"dense (reference-exact) Adam over the whole {rows}x{D4} table: 24 B per element + 4 B slot per row")
    big_b = 1 << 20
    ids2 = torch.randint(0, rows, (big_b,), device=dev, generator=g)
    grads = torch.randn(big_b, D4, **f32)
    wsb = lib.rb200_scatter_workspace_bytes(big_b, rows)
    ws = _lib.workspace(wsb, dev)
    uq = torch.empty(big_b, dtype=torch.int64, device=dev)
    ugr = torch.empty(big_b, D4, **f32)
    nuq = torch.zeros(1, dtype=torch.int32, device=dev)
    ms = timeit(lambda: _lib.check(lib.rb200_scatter_rows(ids2.data_ptr(), grads.data_ptr(), big_b, D4, rows, 0, None, uq.data_ptr(), ugr.data_ptr(),
                                                          nuq.data_ptr(), None, ws.data_ptr(), wsb, sp())))
    n_u = int(nuq.item())
    add("scatter_rows", ms, big_b * (4 * D4 + 8) + n_u * (4 * D4 + 8),
        f"deterministic sorted-segment sum of {big_b} gradient rows into {n_u} unique rows (radix sort of (id, sample) pairs included)")
    # the two phases apart (rb200_scatter_plan needs only the ids: the steps run it on a side stream under the towers; what is left on
    # the critical path is rb200_scatter_apply — rows in, unique rows out)
    _lib.check(lib.rb200_scatter_plan(ids2.data_ptr(), big_b, rows, 0, uq.data_ptr(), nuq.data_ptr(), None, ws.data_ptr(), wsb, sp()))
    ms_apply = timeit(lambda: _lib.check(lib.rb200_scatter_apply(grads.data_ptr(), big_b, D4, rows, None, uq.data_ptr(), ugr.data_ptr(),
                                                                 nuq.data_ptr(), ws.data_ptr(), wsb, sp())))
    add("scatter_apply", ms_apply, big_b * (4 * D4 + 4) + n_u * (4 * D4),
        f"second phase alone (segment sums of {big_b} rows into {n_u} unique rows; the id sort is rb200_scatter_plan, off the critical path)")
    return res


def make_flat_shard(dev, rows: int, seed: int):
    """C5 shard: unit-norm N(0,1) rows generated on the device in slices (a 12.5 M x 64 shard is 3.2 GB)."""
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.empty(rows, 64, dtype=torch.float32, device=dev)
    for r0 in range(0, rows, 1 << 20):
        blk = torch.randn(min(1 << 20, rows - r0), 64, device=dev, generator=g)
        x[r0:r0 + blk.shape[0]] = torch.nn.functional.normalize(blk, dim=-1)
    return x


def bench_flat(args, dev, rows: int = 12_500_000, nqs=(4096, 64), reps: int = 3):
    """BASELINE C5, the work of ONE of its 8 shards: exhaustive inner-product top-500 over 12.5 M x 64 rows (100 M / 8) —
    rb200_flat_search = first exact chunk, then threshold-pruned tcgen05 rounds (more than 128 queries: csrc/flat_filter_tc.cu, one-pass
    TF32 filter + exact re-score of the survivors; at most 128: csrc/flat_stream_tc.cu, tensor-map TMA stream).  The N>1 line adds the
    all-gather + merge across shards."""
    import recommendit_b200 as R
    pk = peaks()
    x = make_flat_shard(dev, rows, 13)
    g = torch.Generator(device=dev).manual_seed(14)
    out = {"rows_per_shard": rows, "k": 500, "D": 64}
    for nq in nqs:
        q = torch.nn.functional.normalize(torch.randn(nq, 64, device=dev, generator=g), dim=-1)
        R.flat_search(q, x, 500)                              # warm-up (kernel attributes, workspace)
        torch.cuda.synchronize(dev)
        ms = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); s, i = R.flat_search(q, x, 500); b.record()
            torch.cuda.synchronize(dev)
            ms.append(a.elapsed_time(b))
        t = float(np.median(ms))
        # sanity on a 1M-row prefix against a library matmul + top-k, judged by the parity rule of the tests
        # (oracle/ivf_oracle.assert_topk_equivalent): ids identical except where scores tie within the fp32 rounding of a
        # 64-term dot product (two engines sum in different orders) — an id may differ only if its score equals the
        # reference's score at that rank, or it sits at the k-th-score boundary
        sub = 1_000_000
        s1, i1 = R.flat_search(q[:64], x[:sub], 500)
        ref = torch.topk(q[:64].double() @ x[:sub].double().T, 500, dim=1)
        same = i1 == ref.indices
        near = (s1.double() - ref.values).abs() <= 4e-7
        kth = ref.values[:, -1:]
        in_ref = (s1.double() >= kth - 4e-7)
        parity_ok = bool(((same | near) & in_ref).all().item())
        flops = 2.0 * nq * rows * 64
        ent = {"ms_per_batch": t, "queries_per_s": nq / t * 1e3, "logical_tflops": flops / (t * 1e-3) / 1e12,
               "prefix_1M_parity": {"ok": parity_ok, "ids_identical_frac": float(same.float().mean()),
                                    "ids_differing_are_score_ties_within": 4e-7, "max_score_diff": float((s1.double() - ref.values).abs().max()),
                                    "rule": "ids identical except exact/rounding-level score ties (north_star); scores within 2e-6"}}
        if nq > 128:
            issued = flops * 9.0 / 8.0                        # 8 K steps of scores + 1 K step that subtracts the threshold
            ent["roofline"] = {"kernel": "flat_filter_tc_kernel (one-pass TF32 filter, N = 128 MMAs, threshold compared on the tensor core; "
                                         "survivors re-scored in fp32 by flat_rescore_kernel)",
                               "bound": "tensor", "unit": "TFLOP/s",
                               "achieved": flops / (t * 1e-3) / 1e12, "issued_tflops": issued / (t * 1e-3) / 1e12,
                               "peak": pk["bf16_tflops"], "frac": flops / (t * 1e-3) / 1e12 / pk["bf16_tflops"],
                               "frac_of_tf32_peak": flops / (t * 1e-3) / 1e12 / (pk["bf16_tflops"] / 2.0),
                               "traffic": ncu_traffic("flat_filter_tc_kernel"),
                               "traffic_note": "per launch of the largest round (6.29 M rows = 1.61 GB of the 12.5 M-row shard): "
                                               "profiles/r02_flat_filter.md",
                               "note": "logical 2·nq·N·D flops of the WHOLE search (all rounds, re-score and selects included in the time) "
                                       "vs the measured bf16 peak; kind::tf32 peaks at half of it (frac_of_tf32_peak)"}
        else:
            by = rows * 64 * 4.0
            ent["roofline"] = {"kernel": "flat_stream_tc_kernel (persistent, tensor-map TMA, one-pass TF32 filter)", "bound": "hbm", "unit": "GB/s",
                               "achieved": by / (t * 1e-3) / 1e9,
                               "peak": pk["hbm_gbs"], "frac": by / (t * 1e-3) / 1e9 / pk["hbm_gbs"], "traffic": ncu_traffic("flat_stream_tc_kernel"),
                               "traffic_note": "per launch of the largest round (8.31 M rows = 2.13 GB): profiles/r02_flat_stream.md",
                               "note": "the shard is read once per batch: N·D·4 bytes over the time of the WHOLE search (four rounds, "
                                       "re-score and selects included)"}
        out[f"nq{nq}"] = ent
    del x
    torch.cuda.empty_cache()
    return out


def bench_flat_cpu(nq: int = 16, rows: int = 1_000_000):
    from oracle import ivf_oracle as V
    rng = np.random.default_rng(13)
    x = V.normalize_rows(rng.standard_normal((rows, 64)).astype(np.float32))
    q = V.normalize_rows(rng.standard_normal((nq, 64)).astype(np.float32))
    V.flat_search_c(q[:2], x[:10000], 500)
    t0 = time.perf_counter()
    V.flat_search_c(q, x, 500)
    dt = time.perf_counter() - t0
    return {"value": nq / dt * rows / 12_500_000, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{nq} queries x {rows} rows with oracle/ivf_oracle.c (heap-based IndexFlatIP restatement, OpenMP over queries), "
                      "scaled linearly in rows to the 12.5 M-row shard"}


def bench_serving(dev, n_req: int = 300):
    """The serving micro-path at the reference's own sizes (BASELINE C1: 6040 users, 3883 catalog items, D = 64, IVFFlat nlist 100 /
    nprobe 10, top-500): per request `model.get_user_embedding(user_id)` + `FAISSIndex.search(vec, 500)` through the drop-in
    classes, numpy in / numpy out, one request at a time (serving/recommender.py:148-156,203); plus all users in one batch."""
    import recommendit_b200 as R
    rng = np.random.default_rng(SEED)
    catalog = np.sort(rng.choice(np.arange(1, N_ITEMS + 1), 3883, replace=False)).astype(np.int64)
    genres = (rng.random((len(catalog), E)) < 0.092).astype(np.float32)
    torch.manual_seed(0)
    model = R.TwoTowerModel(N_USERS, N_ITEMS, D, H, dropout=DROPOUT).to(dev).eval()
    item_emb = model.get_item_embeddings([int(i) for i in catalog], genres, dev)
    idx = R.FAISSIndex(D, 100, 10)
    idx.build_ivf_index(item_emb, [int(i) for i in catalog])
    users = rng.integers(1, N_USERS + 1, n_req + 20)
    lat = []
    for j, u in enumerate(users):
        t0 = time.perf_counter()
        v = model.get_user_embedding(int(u), dev)
        s, ids = idx.search(v, 500)
        dt = time.perf_counter() - t0
        if j >= 20:
            lat.append(dt * 1e3)
    lat = np.sort(np.array(lat))
    graph_req = None
    try:                                                   # the same request as ONE CUDA-graph replay
        rec = R.UserRecommender(model, idx, k=500)
        lat_g = []
        for j, u in enumerate(users):
            t0 = time.perf_counter()
            sg, ig = rec.recommend(int(u))
            dt = time.perf_counter() - t0
            if j >= 20:
                lat_g.append(dt * 1e3)
        lat_g = np.sort(np.array(lat_g))
        graph_req = {"p50": float(lat_g[len(lat_g) // 2]), "p99": float(lat_g[int(len(lat_g) * 0.99)]), "mean": float(lat_g.mean()),
                     "equals_the_two_calls": bool(np.array_equal(ig, ids) and np.array_equal(sg, s)),
                     "api": "UserRecommender.recommend(user_id): id up through a pinned slot, user tower + IVF search captured in "
                            "one CUDA graph, results back through pinned buffers, one synchronisation"}
    except Exception as e:                                 # a secondary line must not take the headline down with it
        graph_req = {"error": str(e)[:300]}
    all_u = torch.arange(1, N_USERS + 1, device=dev)
    with torch.no_grad():
        q = model.user_tower(all_u).cpu().numpy()
    idx.batch_search(q[:64], 500)
    t0 = time.perf_counter()
    bs, bi = idx.batch_search(q, 500)
    tb = time.perf_counter() - t0
    return {"per_request_ms": {"p50": float(lat[len(lat) // 2]), "p99": float(lat[int(len(lat) * 0.99)]), "mean": float(lat.mean())},
            "graph_request_ms": graph_req,
            "requests": len(lat), "results_per_request": int(len(ids)),
            "all_users_batch": {"users": int(N_USERS), "ms": tb * 1e3, "queries_per_s": N_USERS / tb},
            "api": "TwoTowerModel.get_user_embedding(user_id, device) + FAISSIndex.search(vec, 500): numpy in / numpy out, "
                   "wall clock per request, one request at a time",
            "reference": "README.md:42 quotes 6 ms p50 for its whole API request (retrieval + ranking + feature store) on CPU"}


def bench_ivf_cpu(nq_sample: int = 256):
    """CPU arm of C3 on a bounded sample of queries: oracle/ivf_oracle.c (heap-based, OpenMP over queries)."""
    from oracle import ivf_oracle as V
    n, d, nlist, nprobe, k = 1_000_000, 64, 4096, 32, 500
    rng = np.random.default_rng(7)
    cen = V.normalize_rows(rng.standard_normal((nlist, d)).astype(np.float32))
    z = np.minimum((rng.random(n) ** 2 * nlist).astype(np.int64), nlist - 1)
    x = V.normalize_rows(cen[z] + 0.35 * rng.standard_normal((n, d)).astype(np.float32))
    q = V.normalize_rows(x[rng.integers(0, n, nq_sample)] + 0.2 * rng.standard_normal((nq_sample, d)).astype(np.float32))
    off, order = V.build_lists(V.assign(x, cen), nlist)
    lv = np.ascontiguousarray(x[order])
    V.ivf_search_c(q[:8], cen, off, lv, order, nprobe, k)
    t0 = time.perf_counter()
    V.ivf_search_c(q, cen, off, lv, order, nprobe, k)
    dt = time.perf_counter() - t0
    return {"value": nq_sample / dt, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{nq_sample} queries of the C3 workload, oracle/ivf_oracle.c (FAISS-semantics restatement, OpenMP); faiss itself "
                      "is not installed"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--skip-ivf", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-hbm", action="store_true")
    ap.add_argument("--skip-inbatch", action="store_true")
    ap.add_argument("--skip-producer", action="store_true")
    ap.add_argument("--skip-flat", action="store_true")
    ap.add_argument("--skip-c4", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 or args.gpus > 1:
        from bench_sharded import main_sharded
        return main_sharded(args)

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    pk = peaks()
    r = bench_train_single(args, dev)
    K = args.steps
    value = B * K / (r["total_ms"] * 1e-3)
    e2e = B * K / r["pipe_s"]
    # dominant stage and its roofline.  Tower MLP flops per sample: fwd 107 520, bwd 2x.
    stages = r["stages"]
    dom = max(stages, key=stages.get)
    flops = {"towers_fwd": 107520.0 * B, "towers_bwd": 215040.0 * B}
    roof = {"kernel": dom, "bound": "tensor", "unit": "TFLOP/s", "peak": pk["bf16_tflops"], "peak_source": pk["source"],
            "traffic": ncu_traffic(dom), "ms": stages[dom], "share_of_step": stages[dom] / sum(stages.values())}
    if dom in flops:
        roof["achieved"] = flops[dom] / (stages[dom] * 1e-3) / 1e12
        roof["frac"] = roof["achieved"] / pk["bf16_tflops"]
        roof["note"] = ("tcgen05 kind::tf32 kernels in 3xTF32 mode (fp32-grade, 3 MMAs issued per logical MMA) measured against the bf16 "
                        "tensor peak; algorithmic flops = 107 520 (fwd) / 215 040 (bwd) per sample x 8192; the stage is bound by "
                        "operand-staging latency at this batch size, not by the tensor pipe (DESIGN.md §7)")
    else:
        by = {"scatter": (3 * B * (4 * D + 8)) * 2.0, "adam": 24.0 * (N_USERS + N_ITEMS + 2) * D + 24.0 * 35456,
              "loss": 6.0 * B * D * 4, "clip": 4.0 * (3 * B * D + 35456)}[dom]
        roof.update({"bound": "hbm", "unit": "GB/s", "peak": pk["hbm_gbs"], "achieved": by / (stages[dom] * 1e-3) / 1e9})
        roof["frac"] = roof["achieved"] / pk["hbm_gbs"]
    line = {
        "metric": "bpr_train_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": 1, "steps": K, "warmup": args.warmup,
        "ms_per_step": r["total_ms"] / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": dict(C2_CONFIG),
        "config_detail": {"l2": "flushed between timed steps (256 MiB write); per-step CUDA events", "api": "FusedBPRTrainer (CUDA graph)",
                          "tower_mode": "tcgen05 3xTF32 (fp32-grade), activations as the TMEM A operand (csrc/tower_ts.cu)"},
        "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": 4,
                "ms_per_step": r["pipe_s"] / K * 1e3, "mean_loss": r["epoch_loss"],
                "api": "FusedBPRTrainer.train_epoch(pinned host batches): the reference's train_epoch loop "
                       "(train_embeddings.py:170-199), wall clock around the call; every step copies its batch host→device "
                       "(copy stream, overlapped with the previous step) and its loss device→host (pinned array, summed at "
                       "the end); no L2 flush inside the call (each batch arrives over PCIe; tables are 2.6 MB)"},
        "e2e_sync_per_step": {"value": B * K / r["e2e_s"], "unit": "samples/s", "ms_per_step": r["e2e_s"] / K * 1e3,
                              "api": "FusedBPRTrainer.load_packed(pinned batch) + step() + loss.item() per step (host "
                                     "synchronisation every step, L2 flushed between steps), wall clock"},
        "e2e_dropin": {"value": B / (r["dropin_ms"] * 1e-3), "unit": "samples/s", "ms_per_step": r["dropin_ms"],
                       "api": "unchanged reference step body (train_embeddings.py:179-194) on the drop-in TwoTowerModel: "
                              "3 tower calls + bpr_loss + backward + clip_grad_norm_ + torch.optim.Adam + loss.item()"},
        "gpu_launches": r["launches_per_step"] * K, "launches_per_step": r["launches_per_step"],
        "ms_per_step_l2_warm": r["warm_ms"], "stage_ms": stages, "roofline": roof, "clocks": r["clocks"], "final_loss": r["loss"],
    }
    if not args.skip_ivf:
        line["ivf"] = bench_ivf(args, dev)
    if not args.skip_inbatch:
        line["inbatch"] = bench_inbatch(args, dev)
    if not args.skip_producer:
        line["device_producer"] = bench_producer(args, dev)
    if not args.skip_hbm:
        line["hbm_kernels"] = bench_hbm_kernels(dev)
    if not args.skip_flat:
        line["c5_shard"] = bench_flat(args, dev)
    if not args.skip_c4:
        # BASELINE C4 at world 1 with the per-rank batch of the N > 1 lines (their headline): the N = 1 point of the C4 scaling curve
        import bench_sharded as BS
        torch.cuda.empty_cache()
        line["c4"] = BS.bench_c4(min(K, 50), args.warmup, dev, 0, 1)
    if not args.skip_ivf:
        line["serving_c1"] = bench_serving(dev)
    if not args.skip_cpu:
        cores = os.cpu_count() or 1
        v, ms, kind = cpu_step_throughput(20, 3, cores)
        line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": cores, "kind": kind, "ms_per_step": ms,
                                "sample": f"20 full steps of batch {B} after 3 warm-up ({_ref_how(kind)}; torch {torch.__version__} "
                                          f"CPU, {cores} threads, dropout on)"}
        if not args.skip_ivf:
            try:
                line["ivf"]["cpu_baseline"] = bench_ivf_cpu()
            except Exception as e:   # the C oracle is optional test infrastructure
                line["ivf"]["cpu_baseline"] = {"unavailable": str(e)[:200]}
        if not args.skip_flat:
            try:
                line["c5_shard"]["cpu_baseline"] = bench_flat_cpu()
            except Exception as e:
                line["c5_shard"]["cpu_baseline"] = {"unavailable": str(e)[:200]}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
