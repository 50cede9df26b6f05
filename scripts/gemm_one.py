import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import magic_b200
from magic_b200 import ops
M, N, K = 1280, 128, 128
x = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16(); b = torch.randn(N, device="cuda")
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(6):
    ops.gemm(x, K, 1, w, 1, K, out, M, N, K, bias=b)
torch.cuda.synchronize()
