"""Host/device time of one ShardedBPRTrainer step at C4 sizes on ONE GPU (world 1): torch profiler op table."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from recommendit_b200.sharded import ShardedBPRTrainer

dev = torch.device("cuda", 0)
NU, NI, D, B = 2_000_000, 200_000, 128, 8192
tr = ShardedBPRTrainer(NU, NI, D, 128, 18, adam_mode="rows", device=dev, seed=11)
g = torch.Generator(device=dev).manual_seed(0)
def batch():
    return (torch.randint(1, NU + 1, (B,), device=dev, generator=g), torch.randint(1, NI + 1, (B,), device=dev, generator=g),
            (torch.rand(B, 18, device=dev, generator=g) < 0.1).float(), torch.randint(1, NI + 1, (B,), device=dev, generator=g),
            (torch.rand(B, 18, device=dev, generator=g) < 0.1).float())
bs = [batch() for _ in range(8)]
for i in range(5):
    tr.step(*bs[i % 8])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(50):
    tr.step(*bs[i % 8])
torch.cuda.synchronize()
print("ms/step (wall):", (time.perf_counter() - t0) / 50 * 1e3)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(10):
        tr.step(*bs[i % 8])
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=50))
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=25, max_name_column_width=50))
