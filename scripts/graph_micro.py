"""Device-side per-launch time of the hot kernels at the workload's shapes: N back-to-back launches captured in a
CUDA graph (no host launch cost), rotating over a few buffer sets, timed with CUDA events around the replay."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import magic_b200
from magic_b200 import ops, _lib

dev = "cuda"
NREP, NSET = 40, 4


def timeit(name, fn, flops=0.0, nbytes=0.0):
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
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (5 * NREP)
    extra = ""
    if flops:
        extra += f"  {flops / us / 1e6:8.1f} TFLOP/s"
    if nbytes:
        extra += f"  {nbytes / us / 1e3:8.1f} GB/s"
    print(f"{name:42s} {us:8.2f} us{extra}", flush=True)


def gemm_cases(shapes):
    for (M, N, K) in shapes:
        xs = [torch.randn(M, K, device=dev).bfloat16() for _ in range(NSET)]
        w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        b = torch.randn(N, device=dev)
        bb = b.bfloat16()
        outs = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(NSET)]
        pres = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(NSET)]
        dys = [torch.randn(M, N, device=dev).bfloat16() for _ in range(NSET)]
        dxs = [torch.empty(M, K, device=dev, dtype=torch.bfloat16) for _ in range(NSET)]
        gw = torch.zeros(N, K, device=dev)
        fl = 2.0 * M * N * K
        timeit(f"gemm fwd   {M}x{N}x{K} bias", lambda i: ops.gemm(xs[i], K, 1, w, 1, K, outs[i], M, N, K, bias=b), fl)
        timeit(f"gemm fwd   {M}x{N}x{K} bias+gelu+pre",
               lambda i: ops.gemm(xs[i], K, 1, w, 1, K, outs[i], M, N, K, bias=b, act=1, pre_out=pres[i]), fl)
        timeit(f"gemm dgrad {M}x{K}x{N}", lambda i: ops.gemm(dys[i], N, 1, w, K, 1, dxs[i], M, K, N), fl)
        timeit(f"gemm wgrad {N}x{K}x{M} beta1", lambda i: ops.gemm(dys[i], 1, N, xs[i], K, 1, gw, N, K, M, beta=1.0), fl)
        timeit(f"torch linear {M}x{N}x{K}", lambda i: torch.nn.functional.linear(xs[i], w, bb), fl)
        gb = torch.zeros(N, device=dev)
        timeit(f"colsum {M}x{N}",
               lambda i: _lib.call("magic_colsum", dys[i].data_ptr(), gb.data_ptr(), M, N, N, 1, _lib.stream()),
               0, M * N * 2)


def attn_cases(cases):
    for (B, H, Lq, Lk, bias) in cases:
        h = H * 64
        qkv = [torch.randn(B * Lq, 3 * h, device=dev).bfloat16() for _ in range(NSET)]
        kv = [torch.randn(B * Lk, 2 * h, device=dev).bfloat16() for _ in range(NSET)]
        lens = torch.randint(Lk // 2, Lk + 1, (B,), device=dev, dtype=torch.int32)
        dist = torch.rand(B, Lq, Lk, device=dev) * 30 if bias else None
        sw = torch.tensor([0.01], device=dev) if bias else None
        sb = torch.tensor([0.02], device=dev) if bias else None
        fl = 4.0 * B * H * Lq * Lk * 64
        by = (2 * B * Lq + 2 * B * Lk) * h * 2
        for pbar in (False, True):
            def f(i):
                with torch.no_grad():
                    if Lq == Lk:
                        ops.attention(qkv[i], None, 0, h, 2 * h, B, H, Lq, Lk, lens, dist, sw, sb, pbar, 0.1, 3)
                    else:
                        ops.attention(qkv[i], kv[i], 0, 0, h, B, H, Lq, Lk, lens, None, None, None, pbar, 0.1, 3)
            timeit(f"attn fwd B{B} H{H} {Lq}x{Lk} bias{int(bias)} pbar{int(pbar)}", f, fl, by)
        # raw backward call (saved lse from a raw forward call), timed inside a graph like the rest
        q = qkv[0]
        kvt = q if Lq == Lk else kv[0]
        q_off, k_off, v_off = (0, h, 2 * h) if Lq == Lk else (0, 0, h)
        out = torch.empty(B * Lq, h, device=dev, dtype=torch.bfloat16)
        lse = torch.empty(B, H, Lq, device=dev)
        pb = torch.empty(B, Lq, Lk, device=dev)
        seed = ops.seed_tensor(q.device)
        P = lambda t: None if t is None else t.data_ptr()
        _lib.call("magic_attn_fwd", q.data_ptr() + 2 * q_off, kvt.data_ptr() + 2 * k_off, kvt.data_ptr() + 2 * v_off,
                  q.stride(0), kvt.stride(0), kvt.stride(0), out.data_ptr(), lse.data_ptr(), pb.data_ptr(), Lq * Lk, Lk,
                  B, H, Lq, Lk, lens.data_ptr(), P(dist), P(sw), P(sb), 0.125, 1, 0.1, 3, seed.data_ptr(), _lib.stream())
        do = torch.randn_like(out)
        dp = torch.randn_like(pb) * 1e-3
        delta = torch.empty_like(lse)
        dq = torch.empty_like(q)
        dkv = dq if Lq == Lk else torch.empty_like(kvt)
        gs = torch.zeros(2, device=dev)

        def fb(i):
            _lib.call("magic_attn_bwd", q.data_ptr() + 2 * q_off, kvt.data_ptr() + 2 * k_off,
                      kvt.data_ptr() + 2 * v_off, q.stride(0), kvt.stride(0), kvt.stride(0), do.data_ptr(),
                      lse.data_ptr(), dp.data_ptr(), Lq * Lk, Lk, delta.data_ptr(), dq.data_ptr() + 2 * q_off,
                      dkv.data_ptr() + 2 * k_off, dkv.data_ptr() + 2 * v_off, dq.stride(0), dkv.stride(0),
                      dkv.stride(0), gs.data_ptr() if bias else None, B, H, Lq, Lk, lens.data_ptr(), P(dist), P(sw),
                      P(sb), 0.125, 1, 0.1, 3, seed.data_ptr(), _lib.stream())
        timeit(f"attn bwd B{B} H{H} {Lq}x{Lk} bias{int(bias)} pbar1", fb, 2.5 * fl, 2 * by)


def ln_cases(cases):
    for (M, h) in cases:
        xs = [torch.randn(M, h, device=dev).bfloat16() for _ in range(NSET)]
        rs = [torch.randn(M, h, device=dev).bfloat16() for _ in range(NSET)]
        g = torch.ones(h, device=dev)
        b = torch.zeros(h, device=dev)

        def f(i):
            with torch.no_grad():
                ops.layer_norm(xs[i], g, b, 1e-12, res=rs[i], p_in=0.1, salt_in=5)
        timeit(f"ln fwd {M}x{h} res+drop", f, 0, 3 * M * h * 2)
        st = torch.empty(M, 2, device=dev)
        ys = [torch.empty_like(xs[0]) for _ in range(NSET)]
        seed = ops.seed_tensor(torch.device(dev))
        _lib.call("magic_ln_fwd", xs[0].data_ptr(), rs[0].data_ptr(), g.data_ptr(), b.data_ptr(), ys[0].data_ptr(),
                  st.data_ptr(), M, h, 1e-12, 1, 0.1, 5, 0.0, 0, seed.data_ptr(), _lib.stream())
        dys = [torch.randn(M, h, device=dev).bfloat16() for _ in range(NSET)]
        dxs = [torch.empty_like(xs[0]) for _ in range(NSET)]
        drs = [torch.empty_like(xs[0]) for _ in range(NSET)]
        dg, db = torch.zeros(h, device=dev), torch.zeros(h, device=dev)
        timeit(f"ln bwd {M}x{h} res+drop",
               lambda i: _lib.call("magic_ln_bwd", dys[i].data_ptr(), xs[0].data_ptr(), rs[0].data_ptr(), g.data_ptr(),
                                   st.data_ptr(), dxs[i].data_ptr(), drs[i].data_ptr(), dg.data_ptr(), db.data_ptr(),
                                   M, h, 1, 0.1, 5, 0.0, 0, seed.data_ptr(), _lib.stream()), 0, 5 * M * h * 2)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "gemm_s"):
    gemm_cases([(5120, 128, 128), (5120, 384, 128), (5120, 512, 128), (5120, 128, 512), (8064, 128, 768), (1280, 128, 128)])
if which in ("all", "gemm_l"):
    gemm_cases([(5120, 768, 768), (5120, 2304, 768), (5120, 3072, 768), (5120, 768, 3072)])
if which == "gemm_m":  # MAGIC-L / ICoD shapes (B = 32: 2560 text rows, 1184 local rows, 640 graph rows)
    gemm_cases([(2560, 768, 768), (2560, 2304, 768), (2560, 3072, 768), (2560, 768, 3072), (1184, 768, 768), (640, 768, 768)])
if which in ("all", "attn"):
    attn_cases([(64, 2, 80, 80, False), (224, 2, 36, 36, False), (64, 2, 20, 20, True), (64, 2, 20, 80, False),
                (64, 2, 37, 80, False), (64, 2, 80, 37, False), (64, 12, 80, 80, False)])
if which == "attn_l":
    attn_cases([(64, 12, 80, 80, False), (64, 2, 80, 80, False), (128, 12, 160, 160, False), (128, 2, 160, 160, False),
                (320, 12, 36, 36, False), (64, 12, 20, 20, True), (128, 12, 50, 160, False), (64, 12, 37, 80, False),
                (128, 2, 50, 50, True), (1536, 2, 36, 36, False)])
if which in ("all", "ln"):
    ln_cases([(5120, 128), (8064, 128), (5120, 768), (11520, 768)])
