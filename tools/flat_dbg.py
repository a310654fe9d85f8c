"""one small exhaustive search (tools; debugging aid).  usage: flat_dbg.py rows nq"""
import sys
sys.path.insert(0, "/root/repo")
import torch
import recommendit_b200 as R
rows, nq = int(sys.argv[1]), int(sys.argv[2])
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.nn.functional.normalize(torch.randn(rows, 64, device="cuda", generator=g), dim=-1)
q = torch.nn.functional.normalize(torch.randn(nq, 64, device="cuda", generator=g), dim=-1)
try:
    s, i = R.flat_search(q, x, 100)
    ref = torch.topk(q @ x.T, 100, dim=1)
    print("ok", float((i == ref.indices).float().mean()), float((s - ref.values).abs().max()))
except Exception as e:
    print("ERR", e)
