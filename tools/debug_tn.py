import sys; sys.path.insert(0, '.')
import torch
from recommendit_b200.faiss_index import scores_tn
K, M, N = 8, 128, 128
A = torch.ones(K, M); B = torch.ones(K, N)
for dbg in (0, 1, 2, 3):
    C = scores_tn(A.cuda(), B.cuda(), 1 | (dbg << 2)).cpu()
    print("dbg", dbg, "(bit0: A K-major flag, bit1: B K-major flag) unique", torch.unique(C)[:8].tolist())
