"""Launch rb200_scatter_rows on a large batch a few times (for an ncu launch list of the large-batch path)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from recommendit_b200 import _lib

lib = _lib.load()
B, D, rows = 1 << 20, 128, 4_000_000
g = torch.Generator(device="cuda").manual_seed(0)
ids = torch.randint(1, rows, (B,), device="cuda", generator=g)
grads = torch.randn(B, D, device="cuda", generator=g)
uq = torch.empty(B, dtype=torch.int64, device="cuda"); ug = torch.empty(B, D, device="cuda")
nu = torch.zeros(1, dtype=torch.int32, device="cuda")
wsb = lib.rb200_scatter_workspace_bytes(B, rows)
ws = _lib.workspace(wsb, "cuda")
for _ in range(3):
    _lib.check(lib.rb200_scatter_rows(ids.data_ptr(), grads.data_ptr(), B, D, rows, 0, None, uq.data_ptr(), ug.data_ptr(), nu.data_ptr(), None,
                                      ws.data_ptr(), wsb, _lib.stream_ptr()))
torch.cuda.synchronize()
print("n_uniq", int(nu.item()))
