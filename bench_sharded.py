"""bench.py, N > 1: BASELINE config C4 — the two-tower BPR step on row-sharded tables (10 M users × 1 M items, D = 128,
H = 128), 8192 samples per rank per step (weak scaling), NCCL all-to-all for ids / rows / row gradients, all-reduce for
the MLP gradients and the clip/loss scalars.  Launched by torchrun, one rank per GPU."""
from __future__ import annotations

import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist

N_USERS, N_ITEMS, D, H, E, B = 10_000_000, 1_000_000, 128, 128, 18, 8192


def main_sharded(args):
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    from bench import ClockSampler, peaks
    from recommendit_b200 import _lib
    from recommendit_b200.sharded import ShardedBPRTrainer
    lib = _lib.load()
    tr = ShardedBPRTrainer(N_USERS, N_ITEMS, D, H, E, adam_mode="rows", device=dev, seed=11)
    K, W = args.steps, args.warmup
    nb = min(K + W, 16)
    rng = np.random.default_rng(1000 + rank)
    host = []
    for _ in range(nb):
        u = (rng.zipf(1.05, B) - 1) % N_USERS + 1                       # Zipf(1.05) over users
        p, n = rng.integers(1, N_ITEMS + 1, B), rng.integers(1, N_ITEMS + 1, B)
        pg, ng = (rng.random((B, E)) < 0.092).astype(np.float32), (rng.random((B, E)) < 0.092).astype(np.float32)
        host.append(tuple(torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (u, p, pg, n, ng)))
    resident = [tuple(t.to(dev) for t in b) for b in host]
    c0 = lib.rb200_launch_count()
    tr.step(*resident[0])
    launches_per_step = lib.rb200_launch_count() - c0
    for i in range(W):
        tr.step(*resident[(1 + i) % nb])
    sampler = ClockSampler(local) if rank == 0 else None
    # ---- value: device-resident batches; barrier + synchronize on both sides; max over ranks ---------------- #
    dist.barrier(); torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        loss = tr.step(*resident[(1 + W + i) % nb])
    e1.record()
    torch.cuda.synchronize(dev); dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    # ---- e2e: pinned host batch → H2D → step (the step reads the loss back) --------------------------------- #
    dist.barrier(); torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for i in range(K):
        b = tuple(x.to(dev, non_blocking=True) for x in host[(1 + W + i) % nb])
        loss = tr.step(*b)
    torch.cuda.synchronize(dev)
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        pk = peaks()
        value = world * B * K / (total_ms * 1e-3)
        h2d = sum(x.numel() * x.element_size() for x in host[0])
        # HBM-side algorithmic bytes per sample in touched-rows mode: ≈ 96·D + 170 B
        bytes_per_sample = 96 * D + 170
        line = {
            "metric": "bpr_train_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"C4: two-tower BPR step on row-sharded tables, 10M users x 1M items, D=128, H=128, "
                                   f"{B} samples per rank per step (global {world * B}), ids Zipf(1.05) users / uniform items, "
                                   "sampled negatives + bpr_loss, clip 1.0, Adam wd 1e-5 on touched rows",
                       "parallelism": f"row-sharded tables x{world} (id mod world), replicated MLPs; NCCL all-to-all ids/rows/grads, "
                                      "all-reduce MLP grads + scalars", "l2": "tables (5.6 GB) far exceed L2"},
            "e2e": {"value": world * B * K / float(te.item()), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 + 128},
            "gpu_launches": int(launches_per_step) * K, "launches_per_step": int(launches_per_step),
            "roofline": {"kernel": "whole step (HBM side)", "bound": "hbm", "achieved": bytes_per_sample * B / (total_ms / K * 1e-3) / 1e9,
                         "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": bytes_per_sample * B / (total_ms / K * 1e-3) / 1e9 / pk["hbm_gbs"],
                         "traffic": None, "peak_source": pk["source"],
                         "note": "per-GPU; the step is bound by collective latency and host orchestration at this batch size, not HBM"},
            "clocks": clocks, "final_loss": float(loss),
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()
