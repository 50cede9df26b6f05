"""Where does a small tensor-core GEMM spend its life?  CTA 0 stamps clock64() at its phase boundaries
(magic_gemm_set_trace); the GEMM runs as the LAST of a chain of identical launches captured in a CUDA graph (warm
instruction cache, warm descriptors, no host launch cost).  Also prints the per-launch time of the chain and the
floor of an empty kernel in the same kind of chain.

  python scripts/gemm_trace.py            # the MAGIC-S shapes
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import magic_b200
from magic_b200 import ops, _lib

dev = "cuda"
NREP, NSET = 40, 4
NAMES = ["entry", "setup done", "first TMA issued", "first stage landed", "last MMA committed", "accum ready (epi)",
         "tmem ld done", "epi math+staged", "TMA store issued", "stores read smem", "final sync", "dealloc",
         "  (first k-block MMAs issued)", "  (second stage landed)", "  (producer: first empty wait passed)"]


def chain(fn, reps=5):
    for i in range(NSET):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g):
            for i in range(NREP):
                fn(i % NSET)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * NREP)


def main():
    sm_mhz = 1965.0
    us = chain(lambda i: _lib.call("magic_delay", 0, _lib.stream()))
    print(f"empty kernel in a graph chain: {us:.2f} us per launch")
    trace = torch.zeros(32, dtype=torch.int64, device=dev)
    seed = ops.seed_tensor(torch.device(dev, 0))
    shapes = [(5120, 128, 128, "bias"), (5120, 384, 128, "bias"), (5120, 512, 128, "gelu"), (5120, 128, 512, "res"),
              (11520, 128, 128, "bias"), (11520, 512, 128, "gelu"), (768, 50272, 128, "bias"), (5120, 768, 768, "bias"),
              (5120, 768, 3072, "res")]
    for (M, N, K, kind) in shapes:
        xs = [torch.randn(M, K, device=dev).bfloat16() for _ in range(NSET)]
        w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        b = torch.randn(N, device=dev)
        outs = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(NSET)]
        pres = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(NSET)]
        res = [torch.randn(M, N, device=dev).bfloat16() for _ in range(NSET)]

        def fn(i):
            if kind == "gelu":
                ops.gemm(xs[i], K, 1, w, 1, K, outs[i], M, N, K, bias=b, act=1, pre_out=pres[i])
            elif kind == "res":
                ops.gemm(xs[i], K, 1, w, 1, K, outs[i], M, N, K, bias=b, residual=res[i], drop_p=0.1, salt=3, seed=seed)
            else:
                ops.gemm(xs[i], K, 1, w, 1, K, outs[i], M, N, K, bias=b)
        _lib.load().magic_gemm_set_trace(None)
        us = chain(fn)
        _lib.load().magic_gemm_set_trace(trace.data_ptr())
        trace.zero_()
        us_tr = chain(fn, reps=2)
        _lib.load().magic_gemm_set_trace(None)
        t = trace.cpu().tolist()
        print(f"\ngemm {M}x{N}x{K} [{kind}]: {us:.2f} us per launch in a graph chain ({us_tr:.2f} with tracing); "
              f"CTA 0, first tile, cycles since entry (us at {sm_mhz:.0f} MHz):")
        for i, n in enumerate(NAMES):
            if t[i]:
                d = t[i] - t[0]
                print(f"   {n:22s} {d:8d}  {d / sm_mhz:7.2f} us")
        if t[16] and t[26]:
            print(f"   globaltimer entry -> final sync: {(t[26] - t[16]) / 1e3:.2f} us")


if __name__ == "__main__":
    main()
