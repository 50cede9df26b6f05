import os, sys, ctypes
os.environ["MAGIC_TC_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import magic_b200
from magic_b200 import ops, _lib
M, N, K = 1280, 128, 128
x = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16(); b = torch.randn(N, device="cuda")
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(6):
    ops.gemm(x, K, 1, w, 1, K, out, M, N, K, bias=b)
torch.cuda.synchronize()
lib = _lib.load()
lib.magic_tc_debug_buffer.restype = ctypes.c_void_p
p = lib.magic_tc_debug_buffer()
buf = torch.empty(10 * 8, dtype=torch.int64, device="cuda")
ctypes.CDLL("libcudart.so.12").cudaMemcpy(ctypes.c_void_p(buf.data_ptr()), ctypes.c_void_p(p), 10 * 8 * 8, 3)
t = buf.cpu().view(10, 8)
t0 = t[:, 0].min()
print("per-CTA timestamps (ns since first CTA start): start, setup_done, first_tma_landed, mma_done, epi_done, dealloc")
for i in range(10):
    print([int(v - t0) for v in t[i, :6]])
