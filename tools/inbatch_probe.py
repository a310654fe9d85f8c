"""Time rb200_bpr_inbatch (SIMT mode 0 vs tcgen05 modes 1/2) at the BASELINE batch."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from recommendit_b200 import _lib

lib = _lib.load()
B, D = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 64
g = torch.Generator(device="cuda").manual_seed(0)
U = torch.nn.functional.normalize(torch.randn(B, D, device="cuda", generator=g), dim=-1)
I = torch.nn.functional.normalize(torch.randn(B, D, device="cuda", generator=g), dim=-1)
loss = torch.empty(1, device="cuda"); dU = torch.empty_like(U); dI = torch.empty_like(I)
wsb = lib.rb200_bpr_inbatch_workspace_bytes(B, D)
ws = _lib.workspace(wsb, "cuda")
modes = tuple(int(m) for m in sys.argv[2].split(",")) if len(sys.argv) > 2 else (0, 2, 1)
for mode in modes:
    def run():
        _lib.check(lib.rb200_bpr_inbatch(U.data_ptr(), I.data_ptr(), B, D, mode, loss.data_ptr(), dU.data_ptr(), dI.data_ptr(), 1.0,
                                         ws.data_ptr(), wsb, _lib.stream_ptr()))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        run()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"mode {mode}: {ms:.3f} ms  loss {loss.item():.7f}  logical {6.0 * B * B * D / ms / 1e9:.1f} TFLOP/s (6·B²·D)")
