"""C4 sharded step: ms per step and the device-timestamp phase table of a graph replay (tools).  Run under torchrun or alone (world 1)."""
import json, os, sys
sys.path.insert(0, "/root/repo")
import torch, torch.distributed as dist
import bench_sharded as BS

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
r = BS.bench_c4(30, 5, dev, rank, world, with_stages=True)
if rank == 0:
    print(json.dumps({"world": world, "ms_per_step": r["ms_per_step"], "value": r["value"], "launches": r["launches_per_step"],
                      "stage_ms_graph": r["stage_ms_graph"], "sum_graph": sum(r["stage_ms_graph"].values()),
                      "strong": r["c4_strong"]["ms_per_step"] if r["c4_strong"] else None}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
