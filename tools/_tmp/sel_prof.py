import sys, ctypes as C
sys.path.insert(0, "/root/repo")
import argparse, torch, bench
from recommendit_b200 import _lib
lib = _lib.load()
lib.rb200_debug_sel_prof.argtypes = [C.c_void_p, C.c_int]
args = argparse.Namespace(steps=3, warmup=3)
lib.rb200_debug_sel_prof(None, 1)
bench.bench_ivf(args, torch.device("cuda", 0))
torch.cuda.synchronize()
out = (C.c_ulonglong * 16)()
lib.rb200_debug_sel_prof(out, 0)
n = out[15]
names = ["load candidates → smem cache", "min/max", "zero hist + histogram (smem atomics)", "suffix scan → boundary bucket", "compaction of survivors", "pad + sort", "stage resolver + resolve ids + write"]
print("queries", n)
tot = 0
for i, nm in enumerate(names):
    print(f"{nm:44s} {out[i] / max(n,1):9.0f} cycles"); tot += out[i] / max(n, 1)
print("total", tot)
