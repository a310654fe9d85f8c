"""bench.py for BASELINE config C4 and the N > 1 lines (launched by torchrun, one rank per GPU, NCCL).

headline (N > 1)   C4: 10 M users x 1 M items, D = 128, H = 128, tables row-sharded across the ranks (id mod world), 8192 samples
                   per rank per step (weak scaling), NCCL all-to-all for ids / rows / row gradients + all-reduce of the MLP
                   gradients and the {sum g^2, loss} pair — `recommendit_b200.sharded.ShardedBPRTrainer`, padded exchange, the whole
                   step incl. the collectives ONE CUDA-graph replay.  The same function produces the ``c4`` object of the N = 1
                   line (world 1, same per-rank batch) so that a C4 scaling efficiency exists: value(N) / (N · c4.value(1)).
"c4_strong"        the same tables at a FIXED global batch of 65 536.
"parity"           in-line evidence (the world-2 pytest is skipped on the driver's 1-GPU test box): a 1/1000 sub-sampled id
                   space of C4 stepped through the SAME exchange / graph path at world N against the single-process CPU oracle
                   on the concatenated batch, and the merged sharded top-500 against the unsharded search.
"c2_dp"            C2 data-parallel replicas (round 1's N > 1 headline), kept as a secondary object.
"c5"               exhaustive top-500 over a row-sharded database.
"""
from __future__ import annotations

import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist

NU4, NI4, D4, H4, E4 = 10_000_000, 1_000_000, 128, 128, 18
DROPOUT4 = 0.1
CAPACITY_FACTOR = 1.5
EXCHANGE = os.environ.get("RB200_C4_EXCHANGE", "p2p")      # "p2p" (peer memory over NVLink) or "padded" (NCCL all-to-alls)


def _dist_on() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def _timed(step_fn, K, dev):
    """K calls bracketed by barrier + synchronize, device-timed (CUDA events on the launching stream), MAX over ranks → ms"""
    if _dist_on():
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        step_fn(i)
    e1.record()
    torch.cuda.synchronize(dev)
    if _dist_on():
        dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if _dist_on():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _max_over_ranks(x: float, dev) -> float:
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if _dist_on():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def c4_config(world: int, B: int = 8192):
    """the `config` object of the C4 line at `world` ranks (also printed by the reference arm: the driver compares the two)"""
    import math
    from recommendit_b200.sharded import shard_rows
    cap = min(3 * B, int(math.ceil(3 * B / world * CAPACITY_FACTOR)) + 16)
    table_bytes = (shard_rows(NU4 + 1, world, 0) + shard_rows(NI4 + 1, world, 0)) * D4 * 4
    return {"workload": f"C4: 10M users x 1M items, D=128, H=128, 18 genres, tables row-sharded x{world} (id mod world; "
                        f"{table_bytes / 1e9:.2f} GB of rows + 2x Adam moments per rank), {B} samples per rank per step (global "
                        f"{world * B}), ids Zipf(1.05) users / uniform items, sampled negatives + bpr_loss, dropout {DROPOUT4}, "
                        "clip_grad_norm_ 1.0, Adam wd 1e-5 on the touched rows (rows mode: the dense reference-exact update "
                        "would move all 11M rows, 33.8 GB per step — DESIGN.md §5)",
            "parallelism": _parallelism(world, cap)}


def _parallelism(world: int, cap: int) -> str:
    if world == 1:
        return "world 1: same code path, no collectives"
    if EXCHANGE == "p2p":
        return (f"row-sharded x{world}, peer-memory exchange over NVLink/NVSwitch (torch symmetric memory): every rank READS the rows it "
                f"needs straight from the owners' shards and WRITES its row gradients into the owners' receive buckets ({cap} rows per "
                f"rank pair = {CAPACITY_FACTOR}x the mean) — one kernel each; the MLP gradients (273 KB) and {{sum g^2, loss}} are reduced by "
                "one-shot peer reads in rank order; the exchange plan, the row lists and the owners' sort of the received rows run on a "
                "side stream under the gather and the towers; four cross-GPU barriers per step, NO NCCL inside the step, no host "
                "synchronisation: the whole step is ONE CUDA-graph replay")
    return (f"row-sharded x{world}: NCCL all-to-all of ids / rows / row gradients + all-reduce of the MLP gradients "
            f"and of {{sum g^2, loss}}; fixed-capacity exchange buckets ({cap} rows per rank pair = "
            f"{CAPACITY_FACTOR}x the mean), no host synchronisation: the whole step incl. the collectives is ONE "
            "CUDA-graph replay")


def c4_batches(rank: int, B: int, n: int):
    """ids Zipf(1.05) over users, uniform over items; genres Bernoulli(0.092)"""
    rng = np.random.default_rng(1000 + rank)
    out = []
    for _ in range(n):
        u = (rng.zipf(1.05, B) - 1) % NU4 + 1
        p, ng_ = rng.integers(1, NI4 + 1, B), rng.integers(1, NI4 + 1, B)
        pg, ngg = (rng.random((B, E4)) < 0.092).astype(np.float32), (rng.random((B, E4)) < 0.092).astype(np.float32)
        out.append(tuple(np.ascontiguousarray(a) for a in (u, p, pg, ng_, ngg)))
    return out


def bench_c4(K: int, W: int, dev, rank: int, world: int, B: int = 8192, strong_global: int = 65536, with_stages: bool = True):
    """The C4 sharded step at `world` ranks: device-timed value, end-to-end from pinned host batches, per-phase times, the
    strong-scaling point.  Returns a dict (meaningful on every rank; rank 0 prints it)."""
    from bench import ClockSampler, peaks
    from recommendit_b200 import _lib
    from recommendit_b200.sharded import ShardedBPRTrainer
    lib = _lib.load()
    pk = peaks()
    tr = ShardedBPRTrainer(NU4, NI4, D4, H4, E4, adam_mode="rows", device=dev, seed=11, exchange=EXCHANGE, capacity_factor=CAPACITY_FACTOR,
                           use_cuda_graph=True, dropout=DROPOUT4, check_every=0)
    host = c4_batches(rank, B, 8)
    pinned = [tuple(torch.from_numpy(a).pin_memory() for a in b) for b in host]
    resident = [tuple(t.to(dev) for t in b) for b in pinned]
    h2d = sum(t.numel() * t.element_size() for t in pinned[0])

    # launches of OUR kernels in one step, counted on the first (eager) step
    c0 = lib.rb200_launch_count()
    tr.step(*resident[0])
    launches = int(lib.rb200_launch_count() - c0)
    for i in range(1, 4 + W):                      # second eager step, graph capture, replays
        tr.step(*resident[i % 8])
    assert tr._graph is not None
    sampler = ClockSampler(dev.index or 0) if rank == 0 else None
    ms = _timed(lambda i: tr.step(*resident[i % 8]), K, dev)
    clocks = sampler.stop() if sampler else None

    # end to end: every step copies its batch from pinned host memory and reads the loss back (host sync per step)
    if _dist_on():
        dist.barrier()
    torch.cuda.synchronize(dev)
    Ke = min(K, 50)
    t0 = time.perf_counter()
    for i in range(Ke):
        b = tuple(t.to(dev, non_blocking=True) for t in pinned[i % 8])
        loss = tr.step(*b).item()
    e2e_sync_s = _max_over_ranks(time.perf_counter() - t0, dev)
    # the epoch loop of the public API: H2D of batch i+1 on a copy stream under step i, 4-byte loss D2H per step, one host sync at the end
    tr.train_epoch([pinned[i % 8] for i in range(4)])
    if _dist_on():
        dist.barrier()
    torch.cuda.synchronize(dev)
    Kp = max(K, 50)
    t0 = time.perf_counter()
    mean_loss = tr.train_epoch([pinned[i % 8] for i in range(Kp)])
    e2e_s = _max_over_ranks(time.perf_counter() - t0, dev)
    tr.check_exchange()                            # no request exceeded the exchange capacity in any step so far

    stages = None
    if with_stages:
        stages = tr.profile_stages(*resident[0], reps=3)
        if _dist_on():
            keys = sorted(stages)
            t = torch.tensor([stages[k] for k in keys], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            stages = {k: float(v) for k, v in zip(keys, t.tolist())}
    # the same boundaries INSIDE a graph replay (device timestamps): no host launch time in these
    stages_g = None
    if with_stages:
        stages_g = tr.profile_stages_graph(*resident[0])
        if _dist_on():
            keys = sorted(stages_g)
            t = torch.tensor([stages_g[k] for k in keys], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            stages_g = {k: float(v) for k, v in zip(keys, t.tolist())}

    # strong scaling: fixed global batch
    Bs = strong_global // world
    strong = None
    if Bs >= 128:
        host_s = c4_batches(rank + 100, Bs, 4)
        res_s = [tuple(torch.from_numpy(a).to(dev) for a in b) for b in host_s]
        for i in range(4):
            tr.step(*res_s[i % 4])
        Ks = min(K, 30)
        ms_s = _timed(lambda i: tr.step(*res_s[i % 4]), Ks, dev)
        tr.check_exchange()
        strong = {"metric": "bpr_train_samples_per_s", "value": strong_global * Ks / (ms_s * 1e-3), "unit": "samples/s",
                  "ms_per_step": ms_s / Ks, "steps": Ks, "global_batch": strong_global, "per_rank_batch": Bs, "scaling": "strong"}
    cap = tr.capacity(3 * B)
    table_bytes = tr.table.numel() * 4
    tr.close()
    del tr
    torch.cuda.empty_cache()

    value = world * B * K / (ms * 1e-3)
    step_ms = ms / K
    # algorithmic HBM bytes of one rank's step (SURVEY.md §8d, D = 128: 96·D + 170 B per sample, throughput mode)
    alg_bytes = B * (96 * D4 + 170)
    # exchange bytes per rank per step: ids out (8 B), rows back (4·D), row gradients out (4·D) per padded slot, MLP all-reduce
    slots = world * cap
    a2a_bytes = slots * (8 + 2 * 4 * D4) * (world - 1) / max(world, 1)
    mlp_floats = H4 * D4 + H4 + D4 * H4 + D4 + H4 * (D4 + E4) + H4 + D4 * H4 + D4
    tower_flops = 617472.0 * B                                     # fwd + bwd MLP flops per sample at D = 128
    roof = {"kernel": "whole C4 step of one rank (ShardedBPRTrainer, graph replay)", "bound": "hbm", "unit": "GB/s",
            "achieved": alg_bytes / (step_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "peak_source": pk["source"],
            "frac": alg_bytes / (step_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "algorithmic_bytes": alg_bytes,
            "traffic": None, "traffic_note": "per-kernel dram bytes: profiles/r02_c4_step.md",
            "note": "12.5 KB per sample x 8192 samples per rank; the step is a chain of ~40 small kernels and (N > 1) 6 collectives, "
                    "latency-bound at this batch size: the HBM fraction of the whole step is small by construction",
            "exchange": {"all_to_all_bytes_per_rank_per_step": int(a2a_bytes), "mlp_allreduce_bytes": mlp_floats * 4,
                         "nvlink_GBps_if_alone": None}}
    if stages:
        src = stages_g or stages
        tw = src.get("towers_fwd", 0.0) + src.get("towers_bwd", 0.0)
        roof["towers"] = {"kernel": "tower_fwd_ts_kernel<128> + tower_bwd_data_ts_kernel<128> + tower_bwd_weights_ts_kernel<128,*> (tcgen05 3xTF32)",
                          "bound": "tensor", "unit": "TFLOP/s", "ms": tw, "achieved": tower_flops / (tw * 1e-3) / 1e12 if tw else None,
                          "peak": pk["bf16_tflops"], "frac": tower_flops / (tw * 1e-3) / 1e12 / pk["bf16_tflops"] if tw else None,
                          "share_of_step": tw / max(sum(src.values()), 1e-9),
                          "timed": "device timestamps inside a graph replay" if stages_g else "eager phases"}
        comm = stages.get("route_gather_exchange", 0.0) + stages.get("grad_exchange", 0.0) + stages.get("allreduce_norm_clip", 0.0)
        roof["exchange"]["phases_with_collectives_ms_eager"] = comm
    return {
        "metric": "bpr_train_samples_per_s", "value": value, "unit": "samples/s", "ms_per_step": step_ms, "steps": K, "n_gpus": world,
        "config": c4_config(world, B),
        "config_detail": {"l2": "not flushed: every step gathers 24 576 random rows of a table shard far larger than L2",
                          "api": f"ShardedBPRTrainer(exchange='{EXCHANGE}', use_cuda_graph=True).step(device batch)",
                          "tower_mode": "tcgen05 3xTF32 (fp32-grade), activations as the TMEM A operand (csrc/tower_ts.cu)",
                          "exchange_capacity_rows": cap, "table_bytes_per_rank": table_bytes},
        "e2e": {"value": world * B * Kp / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_s / Kp * 1e3, "steps": Kp, "mean_loss": mean_loss,
                "api": "ShardedBPRTrainer.train_epoch(pinned host batches): the reference's train_epoch loop (train_embeddings.py:170-199); "
                       "every step copies its batch (ids + genres) host -> device on a copy stream under the previous step and its loss "
                       "device -> host into a pinned array; one host synchronisation per epoch; wall clock around the call, max over ranks"},
        "e2e_sync_per_step": {"value": world * B * Ke / e2e_sync_s, "unit": "samples/s", "ms_per_step": e2e_sync_s / Ke * 1e3, "steps": Ke,
                              "api": "per step: batch from pinned host memory -> device, ShardedBPRTrainer.step, loss.item() (host "
                                     "synchronisation every step)"},
        "gpu_launches": launches * K, "launches_per_step": launches, "stage_ms_eager": stages, "stage_ms_graph": stages_g, "roofline": roof, "clocks": clocks,
        "final_loss": float(loss), "c4_strong": strong,
    }


# --------------------------------------------------------------------------------------------------------- #
# in-line parity (rank 0 checks against the CPU oracle; every rank takes part in the collectives)
# --------------------------------------------------------------------------------------------------------- #
def parity_c4(dev, rank: int, world: int):
    """1/1000 sub-sampled id space of C4 (10 000 users x 1 000 items, D = 128) through the SAME path as the headline (padded
    exchange, CUDA graph from the third step) against oracle/two_tower_oracle.train_step on the concatenated global batch."""
    from oracle import two_tower_oracle as O          # the checker, never the thing measured
    from recommendit_b200.sharded import ShardedBPRTrainer
    NU, NI, B, lr, steps = 10_000, 1_000, 512, 1e-2, 4
    P = O.init_params(NU, NI, D4, H4, seed=5)
    init = {k: torch.from_numpy(v) for k, v in P.items()}
    tr = ShardedBPRTrainer(NU, NI, D4, H4, E4, adam_mode="dense", device=dev, init=init, lr=lr, exchange=EXCHANGE, capacity_factor=2.0,
                           use_cuda_graph=True, dropout=0.0)

    def batch(r, s):
        rng = np.random.default_rng(10 * s + r)
        return ((rng.zipf(1.05, B) - 1) % (NU + 1), rng.integers(0, NI + 1, B), (rng.random((B, E4)) < .2).astype(np.float32),
                rng.integers(0, NI + 1, B), (rng.random((B, E4)) < .2).astype(np.float32))
    losses = [float(tr.step(*[torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in batch(rank, s)])) for s in range(steps)]
    graph = tr._graph is not None
    tr.check_exchange()
    full = tr.full_state()
    tr.close()
    out = None
    if rank == 0:
        S = O.AdamState()
        ref = []
        for s in range(steps):
            parts = [batch(r, s) for r in range(world)]
            gb = tuple(np.concatenate([p[k] for p in parts]) for k in range(5))
            ref.append(float(O.train_step(P, S, gb, lr=lr)[0]))
        perr = max(float(np.abs(full[k].cpu().numpy() - P[k]).max()) for k in O.PARAM_KEYS)
        lerr = float(np.abs(np.array(losses) - np.array(ref)).max())
        out = {"what": f"C4 at 1/1000 of the id space (10 000 users x 1 000 items, D=128), {B} samples per rank, {steps} steps "
                       f"(2 eager + graph capture + replay) at world {world} vs oracle/two_tower_oracle.train_step on the "
                       "concatenated global batch (= the world-1 result), dense Adam",
               "losses": losses, "oracle_losses": ref, "max_loss_err": lerr, "loss_tol": 2e-5,
               "max_param_err": perr, "param_tol": 0.5 * lr, "graph_replayed": graph,
               "ok": bool(lerr <= 2e-5 and perr <= 0.5 * lr and graph)}
    return out


def parity_c5(dev, rank: int, world: int):
    """merged sharded top-500 against the unsharded search of the concatenated database (ids equal except exact-score ties)"""
    from recommendit_b200.faiss_index import flat_search
    from recommendit_b200.sharded import sharded_flat_search
    rows, nq, k = 50_000, 64, 500
    g = torch.Generator(device=dev).manual_seed(21 + rank)
    x = torch.nn.functional.normalize(torch.randn(rows, 64, device=dev, generator=g), dim=-1)
    gq = torch.Generator(device=dev).manual_seed(20)
    q = torch.nn.functional.normalize(torch.randn(nq, 64, device=dev, generator=gq), dim=-1)
    s, i = sharded_flat_search(q, x, k, rank * rows)
    parts = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(parts, x)
    s1, i1 = flat_search(q, torch.cat(parts), k, 0)
    same_ids = (i == i1)
    # A differing id is acceptable only where the two scores at that rank tie at rounding level (4e-7: bench.py's rule for the C5
    # prefix check).  Exact equality cannot be asked for: a row inside a search's exact prefix (its first 8192 rows) is scored by the
    # 3xTF32 GEMM, a row met in a pruned round by the fp32 re-score of the survivors — the same (query, row) pair can differ in the
    # last bits between the sharded search (rows 50 000 … 58 191 open the second shard's prefix) and the unsharded one.
    tie_ok = same_ids | ((s - s1).abs() <= 4e-7)
    return {"what": f"sharded_flat_search over {world} x {rows} rows vs rb200_flat_search over the concatenated {world * rows} rows, "
                    f"{nq} queries, top-{k}", "ids_equal_frac": float(same_ids.float().mean()),
            "rule": "ids identical except where the scores at that rank tie within 4e-7 (fp32 rounding of a 64-term dot product)",
            "max_score_diff": float((s - s1).abs().max()), "ok": bool(tie_ok.all().item() and float((s - s1).abs().max()) <= 2e-6)}


def bench_ivf_sharded(dev, rank: int, world: int, reps: int = 5):
    """C3 scaled out: the 1 M x 64 IVFFlat database row-sharded over the ranks (all centroids + a slice of every
    list per rank), 4096 replicated queries, per-shard top-500 -> NCCL all-gather -> rb200_topk_merge; checked in-line against
    the unsharded index over the same rows and centroids."""
    import recommendit_b200 as R
    from recommendit_b200.sharded import ShardedIVFIndex
    n, d, nlist, nprobe, k, nq = 1_000_000, 64, 4096, 32, 500, 4096
    g = torch.Generator(device=dev).manual_seed(7)                         # the same database on every rank (C3 generator of bench.py)
    cen = torch.nn.functional.normalize(torch.randn(nlist, d, device=dev, generator=g), dim=-1)
    z = (torch.rand(n, device=dev, generator=g) ** 2 * nlist).long().clamp_(max=nlist - 1)
    x = torch.nn.functional.normalize(cen[z] + 0.35 * torch.randn(n, d, device=dev, generator=g), dim=-1)
    q = torch.nn.functional.normalize(x[torch.randint(0, n, (nq,), device=dev, generator=g)] +
                                      0.2 * torch.randn(nq, d, device=dev, generator=g), dim=-1)
    lo, hi = rank * n // world, (rank + 1) * n // world
    sh = ShardedIVFIndex(d, nlist, nprobe)
    sh.build(x[lo:hi].cpu().numpy(), list(range(lo, hi)), centroids=cen.cpu().numpy())
    s, i = sh.search_device(q, k)
    ms = _timed(lambda _: sh.search_device(q, k), reps, dev) / reps
    full = R.FAISSIndex(d, nlist, nprobe)
    full.build_ivf_index(x.cpu().numpy(), list(range(n)), centroids=cen.cpu().numpy())
    s1, i1 = full.index.search_device(q, k, id_table=full._list_item_ids)
    same = i == i1
    ok = bool((same | (s == s1)).all().item()) and float((s - s1).abs().max()) <= 2e-6
    return {"metric": "ivf_top500_qps", "value": nq / ms * 1e3, "unit": "queries/s", "ms_per_batch": ms,
            "config": {"workload": f"C3 x{world}: IVFFlat nlist=4096 nprobe=32 top-500, 4096 queries per batch, 1M x 64 rows sharded by rows "
                                   f"({(hi - lo)} per rank, every rank holds all centroids and its slice of every list)",
                       "parallelism": "per-shard rb200_ivf_search (list scan + select) + NCCL all-gather of (score, id)[4096, 500] + "
                                      "rb200_topk_merge on every rank"},
            "parity": {"what": "merged sharded result vs the unsharded FAISSIndex over the same rows and centroids",
                       "ids_equal_frac": float(same.float().mean()), "max_score_diff": float((s - s1).abs().max()), "ok": ok}}


def cpu_baseline_c4(steps: int = 2):
    """The reference's step body (stock module from oracle/_ref when shipped, else the ATen port) at C4 WIDTHS on the host cores,
    on a bounded sample: one rank's share of the tables at world 8 (1.25 M users x 125 k items: dense Adam over the whole 11 M rows
    would take minutes per step on CPU) and batch 8192."""
    from bench import reference_stepper
    nu, ni = NU4 // 8, NI4 // 8
    cores = os.cpu_count() or 1
    step, kind = reference_stepper(nu, ni, D4, H4, DROPOUT4, cores)
    rng = np.random.default_rng(3)
    B = 8192
    bs = []
    for _ in range(2):
        u = (rng.zipf(1.05, B) - 1) % nu + 1
        p, n = rng.integers(1, ni + 1, B), rng.integers(1, ni + 1, B)
        pg, ng = (rng.random((B, E4)) < 0.092).astype(np.float32), (rng.random((B, E4)) < 0.092).astype(np.float32)
        bs.append(tuple(torch.from_numpy(np.ascontiguousarray(a)) for a in (u, p, pg, n, ng)))
    step(bs[0])
    t0 = time.perf_counter()
    for i in range(steps):
        step(bs[i % 2])
    dt = (time.perf_counter() - t0) / steps
    return {"value": B / dt, "unit": "samples/s", "cores": cores, "kind": kind, "ms_per_step": dt * 1e3,
            "sample": f"{steps} steps of batch {B} after 1 warm-up at D=128 on 1/8 of the C4 tables ({nu} users x {ni} items; dense "
                      f"torch.optim.Adam over every row as the reference does), torch {torch.__version__} CPU, {cores} threads"}


def main_sharded(args):
    # NCCL may print its version banner on stdout; the contract is ONE JSON line there, so everything before the final
    # print goes to stderr (file-descriptor level, native libraries included).
    import sys
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line = _run(args)
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if line is not None:
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def _run(args):
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    import recommendit_b200 as R
    from bench import B, D, DROPOUT, H, N_ITEMS, N_USERS, peaks, synth_batches
    from recommendit_b200 import _lib
    lib = _lib.load()
    K, W = args.steps, args.warmup
    pk = peaks()

    # ---- headline: C4, row-sharded tables ----------------------------------------------------------------------------- #
    c4 = bench_c4(K, W, dev, rank, world)

    # ---- in-line parity --------------------------------------------------------------------------------------------- #
    parity = {"c4_step_vs_oracle": parity_c4(dev, rank, world), "c5_merge_vs_unsharded": parity_c5(dev, rank, world)}

    # ---- secondary: C2, data-parallel replicas (tables of 2.6 MB: replicating them is the natural layout) --------------- #
    torch.manual_seed(0)
    model = R.TwoTowerModel(N_USERS, N_ITEMS, D, H, dropout=DROPOUT).to(dev).train()
    tr = R.DataParallelBPRTrainer(model, lr=1e-3, weight_decay=1e-5, max_norm=1.0, use_cuda_graph=True, allreduce="p2p")
    dp_multimem = bool(getattr(tr, "_dp_mc", 0))
    K2 = min(K, 50)
    nb = min(K2 + W + 2, 24)
    batches, _ = synth_batches(nb, seed=100 + rank)
    pinned = [tr.pack_host(*b).clone().pin_memory() for b in batches]
    resident = [p.to(dev) for p in pinned]
    tr.load_packed(resident[0])
    c0 = lib.rb200_launch_count()
    tr.step()
    dp_launches = lib.rb200_launch_count() - c0
    for i in range(W + 1):
        tr.load_packed(resident[(1 + i) % nb]); tr.step()

    def dev_step(i):
        tr.load_packed(resident[(2 + W + i) % nb]); tr.step()
    dp_ms = _timed(dev_step, K2, dev)
    dist.barrier(); torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for i in range(K2):
        tr.load_packed(pinned[(2 + W + i) % nb])
        dp_loss = tr.step().item()
    dp_e2e = _max_over_ranks(time.perf_counter() - t0, dev)
    chk = model.user_tower.embedding.weight.detach().double().sum().reshape(1)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    in_sync = bool((lo == hi).item())
    ar_bytes = tr.dp_grads.numel() * 4
    if hasattr(tr, "close"):
        tr.close()
    del tr

    # ---- secondary: C5, exhaustive top-500 over a row-sharded database (12.5 M x 64 per rank; 100 M rows at world 8) -------- #
    from bench import make_flat_shard
    from recommendit_b200.sharded import sharded_flat_search
    rows5, nq5, k5 = 12_500_000, 4096, 500
    x5 = make_flat_shard(dev, rows5, 13 + rank)
    g5 = torch.Generator(device=dev).manual_seed(14)                       # the same queries on every rank
    q5 = torch.nn.functional.normalize(torch.randn(nq5, 64, device=dev, generator=g5), dim=-1)
    s5, i5 = sharded_flat_search(q5, x5, k5, rank * rows5)                  # warm-up
    ms5 = _timed(lambda i: sharded_flat_search(q5, x5, k5, rank * rows5), 3, dev) / 3
    chk5 = i5.double().sum().reshape(1)
    lo5, hi5 = chk5.clone(), chk5.clone()
    dist.all_reduce(lo5, op=dist.ReduceOp.MIN); dist.all_reduce(hi5, op=dist.ReduceOp.MAX)
    shards_hit = int(torch.unique(torch.div(i5, rows5, rounding_mode="floor")).numel())
    del x5

    torch.cuda.empty_cache()
    ivf_sh = bench_ivf_sharded(dev, rank, world)
    parity["ivf_sharded_vs_unsharded"] = ivf_sh["parity"]

    if rank != 0:
        return None
    line = {
        "metric": "bpr_train_samples_per_s", "value": c4["value"], "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": c4["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": c4["config"], "config_detail": c4["config_detail"], "e2e": c4["e2e"], "e2e_sync_per_step": c4["e2e_sync_per_step"],
        "gpu_launches": c4["gpu_launches"],
        "launches_per_step": c4["launches_per_step"], "stage_ms_eager": c4["stage_ms_eager"], "stage_ms_graph": c4["stage_ms_graph"], "roofline": c4["roofline"],
        "clocks": c4["clocks"], "final_loss": c4["final_loss"], "c4_strong": c4["c4_strong"], "parity": parity,
        "scaling_basis": "C4 efficiency at N GPUs = value(N) / (N x c4.value of the N=1 line): bench.py --gpus 1 runs this same function at "
                         "world 1 with the same per-rank batch and reports it as its `c4` object (its headline stays C2, the "
                         "configuration BASELINE.json's metric is quoted on for one GPU)",
        "c2_dp": {"metric": "bpr_train_samples_per_s", "value": world * B * K2 / (dp_ms * 1e-3), "unit": "samples/s",
                  "ms_per_step": dp_ms / K2, "steps": K2,
                  "e2e": {"value": world * B * K2 / dp_e2e, "unit": "samples/s", "h2d_bytes_per_step": pinned[0].numel(), "d2h_bytes_per_step": 4},
                  "launches_per_step": int(dp_launches), "replicas_in_sync": in_sync, "final_loss": float(dp_loss),
                  "config": {"workload": f"C2 x{world}: batch 8192 PER RANK, ML-1M-shape tables replicated, dense Adam, dropout 0.1",
                             "parallelism": f"data parallel x{world}: one all-reduce of the dense gradients per step ({ar_bytes} B) over peer "
                                            "memory between two cross-GPU barriers (torch symmetric memory; no NCCL in the step) — "
                                            + ("reduced inside the NVSwitch (rb200_allreduce_multimem: multimem.ld_reduce / multimem.st)"
                                               if dp_multimem else "two-shot kernel (rb200_allreduce_twoshot)")
                                            + ", captured with both halves of the step in ONE CUDA graph"}},
        "c5": {"metric": "flat_top500_qps", "value": nq5 / ms5 * 1e3, "unit": "queries/s", "ms_per_batch": ms5,
               "logical_tflops_all_gpus": 2.0 * nq5 * rows5 * world * 64 / (ms5 * 1e-3) / 1e12,
               "ranks_agree": bool((lo5 == hi5).item()), "shards_in_result": shards_hit,
               "config": {"workload": f"C5: exhaustive inner-product top-{k5} over {rows5 * world / 1e6:.1f} M x 64 fp32 rows "
                                      f"sharded over {world} GPUs ({rows5 / 1e6:.1f} M rows each), {nq5} queries per batch",
                          "parallelism": "per-shard rb200_flat_search (threshold-pruned tcgen05 scan) + NCCL all-gather of the "
                                         "(score, id)[nq, 500] lists + rb200_topk_merge on every rank"}},
    }
    line["ivf_sharded"] = ivf_sh
    if not args.skip_cpu:
        try:
            line["cpu_baseline"] = cpu_baseline_c4()
        except Exception as e:
            line["cpu_baseline"] = {"unavailable": str(e)[:200]}
    return line
