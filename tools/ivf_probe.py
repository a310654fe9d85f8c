"""C3 IVF search only (bench.bench_ivf) — for ncu launch lists of the search kernels."""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import argparse
import torch
import bench

args = argparse.Namespace(steps=int(sys.argv[1]) if len(sys.argv) > 1 else 3, warmup=3)
out = bench.bench_ivf(args, torch.device("cuda", 0))
print(json.dumps({k: out[k] for k in ("value", "ms_per_batch", "e2e")}))
