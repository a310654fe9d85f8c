"""Probe (tools): torch symmetric memory on this box — rendezvous, peer pointers in OUR kernels, barrier in a CUDA graph.
torchrun --nproc-per-node 2 tools/symm_probe.py"""
import os, sys, time
sys.path.insert(0, "/root/repo")
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
from recommendit_b200 import _lib
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
def say(*a): print(f"[r{rank} {time.time() % 1000:.2f}]", *a, flush=True)
lib = _lib.load()
R, D = 100_000, 128
t = symm_mem.empty(R * D, dtype=torch.float32, device=dev)
say("empty ok")
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
say("rendezvous ok", [hex(p) for p in hdl.buffer_ptrs], "multicast ptr", hdl.multicast_ptr)
table = t.view(R, D)
table.copy_(torch.arange(R, device=dev, dtype=torch.float32)[:, None] + 1000.0 * rank)
hdl.barrier(channel=0, timeout_ms=20000)
say("barrier ok")
peer = (rank + 1) % world
pt = hdl.get_buffer(peer, (R, D), torch.float32)
say("peer row 5:", float(pt[5, 0]))
# our gather kernel reading the PEER table through its raw pointer
ids = torch.randint(0, R, (24576,), device=dev)
out = torch.empty(24576, D, device=dev)
_lib.check(lib.rb200_gather_rows(hdl.buffer_ptrs[peer], ids.data_ptr(), ids.numel(), D, R, out.data_ptr(), _lib.stream_ptr()))
torch.cuda.synchronize()
exp = ids.float() + 1000.0 * peer
say("p2p gather ok:", bool((out[:, 0] == exp).all()), bool((out[:, 127] == exp).all()))
# timing + graph capture of barrier + gather
for _ in range(3):
    hdl.barrier(channel=0, timeout_ms=20000)
    _lib.check(lib.rb200_gather_rows(hdl.buffer_ptrs[peer], ids.data_ptr(), ids.numel(), D, R, out.data_ptr(), _lib.stream_ptr()))
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    hdl.barrier(channel=0, timeout_ms=20000)
    _lib.check(lib.rb200_gather_rows(hdl.buffer_ptrs[peer], ids.data_ptr(), ids.numel(), D, R, out.data_ptr(), _lib.stream_ptr()))
    hdl.barrier(channel=1, timeout_ms=20000)
say("captured")
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dist.barrier(); torch.cuda.synchronize()
a.record()
for _ in range(50):
    g.replay()
b.record(); torch.cuda.synchronize()
say("graph: 2 barriers + p2p gather of 12.6 MB: %.1f us per replay" % (a.elapsed_time(b) / 50 * 1e3))
a.record()
for _ in range(50):
    _lib.check(lib.rb200_gather_rows(hdl.buffer_ptrs[peer], ids.data_ptr(), ids.numel(), D, R, out.data_ptr(), _lib.stream_ptr()))
b.record(); torch.cuda.synchronize()
say("p2p gather alone: %.1f us (%.0f GB/s)" % (a.elapsed_time(b) / 50 * 1e3, 24576 * 512 / (a.elapsed_time(b) / 50 * 1e-3) / 1e9))
a.record()
for _ in range(50):
    _lib.check(lib.rb200_gather_rows(table.data_ptr(), ids.data_ptr(), ids.numel(), D, R, out.data_ptr(), _lib.stream_ptr()))
b.record(); torch.cuda.synchronize()
say("local gather alone: %.1f us" % (a.elapsed_time(b) / 50 * 1e3))
del g
dist.barrier(); dist.destroy_process_group()
say("done")
