import sys
sys.path.insert(0, "/root/repo")
import argparse, torch, bench
args = argparse.Namespace(steps=100, warmup=5)
for p in (0.1, 0.0):
    bench.DROPOUT = p
    r = bench.bench_train_single(args, torch.device("cuda", 0))
    print("dropout", p, "ms/step", r["total_ms"] / 100, r["stages"])
