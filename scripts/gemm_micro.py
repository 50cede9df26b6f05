"""Micro-benchmark of magic_gemm shapes (back-to-back launches, CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import magic_b200
from magic_b200 import ops, _lib

dev = "cuda"
shapes = [(1280, 128, 128), (5120, 128, 128), (5120, 384, 128), (5120, 512, 128), (5120, 128, 512), (11520, 128, 768), (768, 50265, 128), (5120, 768, 768), (5120, 2304, 768), (5120, 3072, 768), (5120, 768, 3072), (8192, 8192, 8192)]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for (M, N, K) in shapes:
    x = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm(x, K, 1, w, 1, K, out, M, N, K, bias=b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.gemm(x, K, 1, w, 1, K, out, M, N, K, bias=b)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    ref = torch.nn.functional.linear(x.float(), w.float(), b)
    err = ((out.float() - ref).norm() / ref.norm()).item()
    print(f"fwd  {M}x{N}x{K}: {us:8.1f} us  {2.0*M*N*K/us/1e6:8.2f} TFLOP/s  rel_err {err:.2e}")
    bb = b.bfloat16()
    for _ in range(3):
        torch.nn.functional.linear(x, w, bb)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        torch.nn.functional.linear(x, w, bb)
    e1.record()
    torch.cuda.synchronize()
    print(f"     cublas(torch) same shape: {e0.elapsed_time(e1) * 1e3 / reps:8.1f} us")
    if M * N <= 1 << 24:
        # wgrad-shaped: C[N,K] = dy^T x  (both MN-major), fp32 out
        dy = torch.randn(M, N, device=dev).bfloat16()
        gw = torch.zeros(N, K, device=dev)
        for _ in range(2):
            ops.gemm(dy, 1, N, x, K, 1, gw, N, K, M)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            ops.gemm(dy, 1, N, x, K, 1, gw, N, K, M)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        ref = dy.float().t() @ x.float()
        err = ((gw - ref).norm() / ref.norm()).item()
        print(f"wgrad {N}x{K}x{M}: {us:8.1f} us  {2.0*M*N*K/us/1e6:8.2f} TFLOP/s  rel_err {err:.2e}")
