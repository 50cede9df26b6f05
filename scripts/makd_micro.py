"""Device-side time of the fused MAKD loss kernels at the BASELINE configs[2] sizes (teacher h=768, B=64, L=80, T=5,
G=20): every MSE segment of a distillation step in ONE launch, and the logit KL of the MLM head [768, 50265].
Launches are captured in a CUDA graph rotating over buffer sets larger than L2, timed with CUDA events; achieved
GB/s = algorithmic bytes (each S/T element once; backward also writes dS) / time, against MEASURED_PEAKS.json.

  python scripts/makd_micro.py [bf16|f32]
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import magic_b200
from magic_b200 import ops, _lib

dev = "cuda"
NSET, NREP = 3, 12


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0


def timeit(fn):
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
    return e0.elapsed_time(e1) * 1e3 / (5 * NREP)


def main():
    dtype = torch.float32 if "f32" in sys.argv else torch.bfloat16
    B, L, T, G, V, h = 64, 80, 5, 20, 37, 768
    hid = [(B, L * h), (B * T, 36 * h), (B * T, h), (B, G * h), (B, V * h)]
    maps = [(B, L * L)] * 6 + [(B * T, 36 * 36)] * 2 + [(B, G * G), (B, G * L)] * 3 + [(B, V * V), (B, V * L)] * 3
    sets = []
    for _ in range(NSET):
        pairs = []
        for (r, c) in hid:
            pairs.append(dict(s=torch.randn(r, c, device=dev).to(dtype), t=torch.randn(r, c, device=dev).to(dtype),
                              w=torch.rand(r, device=dev) if r == B else None, rows=r, inner=c, s_rs=c, t_rs=c,
                              scale=1.0 / (r * c)))
        for (r, c) in maps:
            pairs.append(dict(s=torch.rand(r, c, device=dev), t=torch.rand(r, c, device=dev),
                              w=torch.rand(r, device=dev) if r == B else None, rows=r, inner=c, s_rs=c, t_rs=c,
                              scale=1.0 / (r * c)))
        for p in pairs:
            p["ds"] = torch.empty_like(p["s"])
        sets.append(pairs)
    n = len(sets[0])
    by = sum(p["s"].numel() * p["s"].element_size() + p["t"].numel() * p["t"].element_size() for p in sets[0])
    by_ds = sum(p["s"].numel() * p["s"].element_size() for p in sets[0])
    loss = torch.empty(_lib.MAKD_MAX_SEGS + 1, device=dev)
    gtot = torch.ones(1, device=dev)
    segs_f = [ops._mk_segs(ps, False) for ps in sets]
    segs_b = [ops._mk_segs(ps, True) for ps in sets]
    pk = peak()
    if "once" in sys.argv:  # profiling aid: every kernel twice, eagerly, nothing else (ncu -k regex:makd -c 8)
        R, C, ld = 768, 50265, 50272
        ss = torch.randn(R, ld, device=dev).to(dtype)[:, :C]
        ts = torch.randn(R, ld, device=dev).to(dtype)[:, :C]
        ds = torch.empty(R, ld, device=dev, dtype=dtype)[:, :C]
        w = torch.rand(R, device=dev)
        stats, kl, g1 = torch.empty(R, 2, device=dev), torch.empty(1, device=dev), torch.ones(1, device=dev)
        for i in range(2):
            _lib.call("magic_makd_mse_fwd", segs_f[i], n, loss.data_ptr(), _lib.stream())
            _lib.call("magic_makd_mse_bwd", segs_b[i], n, None, gtot.data_ptr(), _lib.stream())
            _lib.call("magic_makd_kl_fwd", ss.data_ptr(), ts.data_ptr(), R, C, ld, 2.0, w.data_ptr(), 4.0 / R, None,
                      stats.data_ptr(), kl.data_ptr(), _lib.dt(ss), _lib.stream())
            _lib.call("magic_makd_kl_bwd", ss.data_ptr(), ts.data_ptr(), ds.data_ptr(), R, C, ld, 2.0, w.data_ptr(),
                      4.0 / R, None, stats.data_ptr(), g1.data_ptr(), _lib.dt(ss), _lib.stream())
        torch.cuda.synchronize()
        return
    us = timeit(lambda i: _lib.call("magic_makd_mse_fwd", segs_f[i], n, loss.data_ptr(), _lib.stream()))
    print(f"makd_mse fwd  {n} segments {by / 1e6:7.1f} MB: {us:7.2f} us  {by / us / 1e3:7.1f} GB/s  "
          f"{by / us / 1e3 / pk:.3f} of {pk:.0f}")
    us = timeit(lambda i: _lib.call("magic_makd_mse_bwd", segs_b[i], n, None, gtot.data_ptr(), _lib.stream()))
    print(f"makd_mse bwd  {n} segments {(by + by_ds) / 1e6:7.1f} MB: {us:7.2f} us  {(by + by_ds) / us / 1e3:7.1f} GB/s  "
          f"{(by + by_ds) / us / 1e3 / pk:.3f} of {pk:.0f}")
    # correctness of the fused launch against a plain torch evaluation of the same sum
    _lib.call("magic_makd_mse_fwd", segs_f[0], n, loss.data_ptr(), _lib.stream())
    ref = sum(((p["s"].float() - p["t"].float()) ** 2 * (p["w"][:, None] if p["w"] is not None else 1.0)).sum().double()
              * p["scale"] for p in sets[0])
    got = loss[_lib.MAKD_MAX_SEGS].item()
    print(f"   total {got:.7f} vs torch {ref.item():.7f}  rel {abs(got - ref.item()) / abs(ref.item()):.2e}")

    R, C, ld = 768, 50265, 50272
    ss = [torch.randn(R, ld, device=dev).to(dtype)[:, :C] for _ in range(NSET)]
    ts = [torch.randn(R, ld, device=dev).to(dtype)[:, :C] for _ in range(NSET)]
    ds = [torch.empty(R, ld, device=dev, dtype=dtype)[:, :C] for _ in range(NSET)]
    w = torch.rand(R, device=dev)
    stats = torch.empty(R, 2, device=dev)
    kl = torch.empty(1, device=dev)
    g1 = torch.ones(1, device=dev)
    dtc = _lib.dt(ss[0])
    esz = ss[0].element_size()
    us = timeit(lambda i: _lib.call("magic_makd_kl_fwd", ss[i].data_ptr(), ts[i].data_ptr(), R, C, ld, 2.0,
                                    w.data_ptr(), 4.0 / R, None, stats.data_ptr(), kl.data_ptr(), dtc, _lib.stream()))
    by = 2 * R * C * esz
    print(f"makd_kl fwd  [{R},{C}] {by / 1e6:7.1f} MB: {us:7.2f} us  {by / us / 1e3:7.1f} GB/s  {by / us / 1e3 / pk:.3f}")
    us = timeit(lambda i: _lib.call("magic_makd_kl_bwd", ss[i].data_ptr(), ts[i].data_ptr(), ds[i].data_ptr(), R, C, ld,
                                    2.0, w.data_ptr(), 4.0 / R, None, stats.data_ptr(), g1.data_ptr(), dtc,
                                    _lib.stream()))
    by = 3 * R * C * esz
    print(f"makd_kl bwd  [{R},{C}] {by / 1e6:7.1f} MB: {us:7.2f} us  {by / us / 1e3:7.1f} GB/s  {by / us / 1e3 / pk:.3f}")
    _lib.call("magic_makd_kl_fwd", ss[0].data_ptr(), ts[0].data_ptr(), R, C, ld, 2.0, w.data_ptr(), 4.0 / R, None,
              stats.data_ptr(), kl.data_ptr(), dtc, _lib.stream())
    s32, t32 = ss[0].float() / 2.0, ts[0].float() / 2.0
    ref = (torch.softmax(t32, 1) * (torch.log_softmax(t32, 1) - torch.log_softmax(s32, 1))).sum(1).double()
    ref = (ref * w.double()).sum().item() * 4.0 / R
    print(f"   kl {kl.item():.7f} vs torch {ref:.7f}  rel {abs(kl.item() - ref) / abs(ref):.2e}")


if __name__ == "__main__":
    main()
