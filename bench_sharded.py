"""bench.py, N > 1 (launched by torchrun, one rank per GPU, NCCL).

headline  C2 weak scaling: the same two-tower BPR step as N = 1 (batch 8192 per rank, ML-1M-shape tables, dense Adam,
          dropout on), replicas kept identical by ONE all-reduce of the dense gradients per step
          (recommendit_b200.DataParallelBPRTrainer).  The tables are 2.6 MB — replicating them is the natural layout.
"c4"      BASELINE config C4 as a secondary object: 10 M users x 1 M items, D = 128, tables row-sharded across the ranks
          (id mod world), NCCL all-to-all for ids / rows / row gradients (recommendit_b200.sharded.ShardedBPRTrainer).
"""
from __future__ import annotations

import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist


def _timed(step_fn, K, dev):
    dist.barrier(); torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        step_fn(i)
    e1.record()
    torch.cuda.synchronize(dev); dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


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
    from bench import B, D, DROPOUT, H, N_ITEMS, N_USERS, ClockSampler, peaks, synth_batches
    from recommendit_b200 import _lib
    from recommendit_b200.sharded import ShardedBPRTrainer
    lib = _lib.load()
    K, W = args.steps, args.warmup
    pk = peaks()

    # ---- headline: C2, data-parallel replicas ------------------------------------------------------------------- #
    torch.manual_seed(0)
    model = R.TwoTowerModel(N_USERS, N_ITEMS, D, H, dropout=DROPOUT).to(dev).train()
    tr = R.DataParallelBPRTrainer(model, lr=1e-3, weight_decay=1e-5, max_norm=1.0, use_cuda_graph=True)
    nb = min(K + W + 2, 24)
    batches, _ = synth_batches(nb, seed=100 + rank)
    pinned = [tr.pack_host(*b).clone().pin_memory() for b in batches]
    resident = [p.to(dev) for p in pinned]
    tr.load_packed(resident[0])
    c0 = lib.rb200_launch_count()
    tr.step()
    launches_per_step = lib.rb200_launch_count() - c0
    for i in range(W + 1):
        tr.load_packed(resident[(1 + i) % nb]); tr.step()
    sampler = ClockSampler(local) if rank == 0 else None

    def dev_step(i):
        tr.load_packed(resident[(2 + W + i) % nb]); tr.step()
    total_ms = _timed(dev_step, K, dev)

    dist.barrier(); torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for i in range(K):
        tr.load_packed(pinned[(2 + W + i) % nb])
        loss = tr.step().item()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if sampler else None
    # replicas must still be identical
    chk = model.user_tower.embedding.weight.detach().double().sum().reshape(1)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    in_sync = bool((lo == hi).item())

    # ---- secondary: C4, row-sharded tables ------------------------------------------------------------------------ #
    NU4, NI4, D4 = 10_000_000, 1_000_000, 128
    st4 = ShardedBPRTrainer(NU4, NI4, D4, 128, 18, adam_mode="rows", device=dev, seed=11, exchange="padded", capacity_factor=2.0,
                            use_cuda_graph=True)
    rng = np.random.default_rng(1000 + rank)
    res4 = []
    for _ in range(8):
        u = (rng.zipf(1.05, B) - 1) % NU4 + 1                           # Zipf(1.05) over users
        p, n = rng.integers(1, NI4 + 1, B), rng.integers(1, NI4 + 1, B)
        pg, ng = (rng.random((B, 18)) < 0.092).astype(np.float32), (rng.random((B, 18)) < 0.092).astype(np.float32)
        res4.append(tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (u, p, pg, n, ng)))
    for i in range(4):                                   # 2 eager steps, graph capture, one replay
        st4.step(*res4[i % 8])
    K4 = min(K, 50)
    ms4 = _timed(lambda i: st4.step(*res4[i % 8]), K4, dev)
    st4.check_exchange()                                 # no request exceeded the exchange capacity

    cap4 = st4.capacity(3 * B)
    st4.close(); del st4                                 # the captured step holds NCCL nodes: release it before the process group goes
    torch.cuda.empty_cache()

    # ---- secondary: C5, exhaustive top-500 over a row-sharded database (12.5 M x 64 per rank; 100 M rows at world 8) -------- #
    from bench import make_flat_shard
    from recommendit_b200.sharded import sharded_flat_search
    rows5, nq5, k5 = 12_500_000, 4096, 500
    x5 = make_flat_shard(dev, rows5, 13 + rank)
    g5 = torch.Generator(device=dev).manual_seed(14)                       # the same queries on every rank
    q5 = torch.nn.functional.normalize(torch.randn(nq5, 64, device=dev, generator=g5), dim=-1)
    s5, i5 = sharded_flat_search(q5, x5, k5, rank * rows5)                  # warm-up
    ms5 = _timed(lambda i: sharded_flat_search(q5, x5, k5, rank * rows5), 3, dev) / 3
    # every rank must hold the same merged result; its ids must come from all shards
    chk5 = i5.double().sum().reshape(1)
    lo5, hi5 = chk5.clone(), chk5.clone()
    dist.all_reduce(lo5, op=dist.ReduceOp.MIN); dist.all_reduce(hi5, op=dist.ReduceOp.MAX)
    shards_hit = int(torch.unique(torch.div(i5, rows5, rounding_mode="floor")).numel())
    del x5

    if rank == 0:
        value = world * B * K / (total_ms * 1e-3)
        ar_bytes = tr.dp_grads.numel() * 4
        line = {
            "metric": "bpr_train_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"C2 x{world}: two-tower BPR training step, batch 8192 PER RANK (global {world * B}), ML-1M-shape "
                                   "tables 6041x64 / 3953x64 replicated, H=128, sampled negatives + bpr_loss, dropout 0.1, "
                                   "clip_grad_norm_ 1.0, Adam wd 1e-5 (dense)",
                       "parallelism": f"data parallel x{world}: one NCCL all-reduce of the dense gradients per step "
                                      f"({ar_bytes} B: MLPs + both tables) between the two halves of the fused step",
                       "l2": "not flushed (N>1 loop is timed as one region); tables are L2-resident by nature at this size",
                       "api": "DataParallelBPRTrainer (2 CUDA graphs + all-reduce per step)", "replicas_in_sync": in_sync},
            "e2e": {"value": world * B * K / float(te.item()), "unit": "samples/s", "h2d_bytes_per_step": pinned[0].numel(),
                    "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches_per_step) * K, "launches_per_step": int(launches_per_step),
            "roofline": {"kernel": "NCCL all-reduce of the dense gradients (the only data-path collective)", "bound": "hbm",
                         "achieved": None, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": None, "traffic": None,
                         "note": "single-GPU kernels are unchanged from N=1 (see the N=1 line for their roofline)"},
            "c4": {"metric": "bpr_train_samples_per_s", "value": world * B * K4 / (ms4 * 1e-3), "unit": "samples/s",
                   "ms_per_step": ms4 / K4, "steps": K4,
                   "config": {"workload": f"C4: row-sharded tables x{world} (id mod world), 10M users x 1M items, D=128, H=128, "
                                          f"{B} samples per rank per step, ids Zipf(1.05) users / uniform items, Adam on touched rows",
                              "parallelism": "NCCL all-to-all ids/rows/row-gradients + all-reduce MLP grads and scalars; fixed-capacity "
                                             f"exchange buffers ({cap4} rows per rank pair, 2x the mean), no host "
                                             "synchronisation: the whole step incl. the collectives is ONE CUDA graph replay"}},
            "c5": {"metric": "flat_top500_qps", "value": nq5 / ms5 * 1e3, "unit": "queries/s", "ms_per_batch": ms5,
                   "logical_tflops_all_gpus": 2.0 * nq5 * rows5 * world * 64 / (ms5 * 1e-3) / 1e12,
                   "ranks_agree": bool((lo5 == hi5).item()), "shards_in_result": shards_hit,
                   "config": {"workload": f"C5: exhaustive inner-product top-{k5} over {rows5 * world / 1e6:.1f} M x 64 fp32 rows "
                                          f"sharded over {world} GPUs ({rows5 / 1e6:.1f} M rows each), {nq5} queries per batch",
                              "parallelism": "per-shard rb200_flat_search (threshold-pruned tcgen05 scan) + NCCL all-gather of the "
                                             "(score, id)[nq, 500] lists + rb200_topk_merge on every rank"}},
            "clocks": clocks, "final_loss": float(loss),
        }
        return line
    return None
