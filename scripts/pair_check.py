"""Quick stand-alone check + timing of the CTA-pair GEMM (run under `timeout`: a protocol bug shows up as a hang)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import magic_b200
from magic_b200 import ops

dev = "cuda"
SHAPES = [(512, 256, 256), (5120, 768, 768), (5120, 768, 3072), (5120, 3072, 768), (5120, 2304, 768),
          (11520, 768, 768), (8192, 8192, 8192)]
if len(sys.argv) > 1 and sys.argv[1] == "small":
    SHAPES = [(640, 768, 768), (1184, 768, 768), (1280, 768, 768), (2368, 768, 768), (2560, 768, 768), (2560, 2304, 768),
              (2560, 3072, 768), (2560, 768, 3072), (1184, 3072, 768), (1184, 768, 3072), (5760, 768, 768),
              (5120, 256, 768), (5120, 512, 512), (5120, 384, 384), (5120, 1536, 384), (5120, 384, 1536),
              (5120, 256, 256), (5120, 1024, 256), (5120, 256, 1024), (11520, 512, 128)]
if len(sys.argv) > 1 and sys.argv[1] == "pair_only":
    SHAPES = [(5120, 768, 3072), (5120, 3072, 768), (11520, 768, 768), (8192, 8192, 8192)]
for (M, N, K) in SHAPES:
    x = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = torch.randn(N, device=dev)
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    ops.gemm(x, K, 1, w, 1, K, out, M, N, K, bias=b)
    torch.cuda.synchronize()
    ref = F.linear(x.float(), w.float(), b)
    err = ((out.float() - ref).norm() / ref.norm()).item()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g):
            for _ in range(10):
                ops.gemm(x, K, 1, w, 1, K, out, M, N, K, bias=b)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 50
    print(f"{M}x{N}x{K}: rel_err {err:.2e}  {us:8.2f} us  {2.0 * M * N * K / us / 1e6:8.1f} TFLOP/s  "
          f"(MAGIC_TC_PAIR={os.environ.get('MAGIC_TC_PAIR', '1')})", flush=True)
