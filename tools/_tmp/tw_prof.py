import sys, ctypes as C
sys.path.insert(0, "/root/repo")
import argparse, torch, bench
from recommendit_b200 import _lib
lib = _lib.load()
lib.rb200_debug_tw_prof.argtypes = [C.c_void_p, C.c_int]
args = argparse.Namespace(steps=50, warmup=5)
lib.rb200_debug_tw_prof(None, 1)
r = bench.bench_train_single(args, torch.device("cuda", 0))
torch.cuda.synchronize()
out = (C.c_ulonglong * 32)()
lib.rb200_debug_tw_prof(out, 0)
n = out[15]
names = ["prologue (tmem alloc, bar init, sync)", "ids load + sync", "gather X + split + store", "wait W1 image + sync", "GEMM1 issue + wait",
         "epilogue 1 (relu/dropout/hid store/stage)", "wait W2 image + sync", "GEMM2 issue + wait", "epilogue 2 (normalise, store)", "teardown"]
print("tiles", n, "ms/step", r["total_ms"] / 50)
tot = 0
for i, nm in enumerate(names):
    print(f"{nm:44s} {out[i] / max(n,1):9.0f} cycles"); tot += out[i] / max(n, 1)
print("total", tot)
