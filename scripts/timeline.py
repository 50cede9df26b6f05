"""Kernel timeline of the graph-replayed step (no nsys in the image): torch.profiler (CUPTI) around a few steps,
dumped as CSV (task step, kernel, stream, start us, duration us) into gpurun_out/ and summarised per step:
wall time, per-stream busy time, idle gaps on the union of all streams.

  python scripts/timeline.py [--workload magic_s_pretrain_b64] [--steps 4] [--out gpurun_out/timeline.csv]
"""
import argparse
import csv
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="magic_s_pretrain_b64")
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--graphs", type=int, default=1)
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "timeline.csv"))
    ap.add_argument("--lookahead", type=int, default=1)
    ap.add_argument("--sync", type=int, default=0, help="synchronize after every step (isolates the steps)")
    args = ap.parse_args()
    import magic_b200
    from magic_b200 import ops
    from magic_b200.graph_index import batch_to_device
    from magic_b200.train_step import PretrainStepper
    dev = torch.device("cuda", 0)
    w = bench.WORKLOADS[args.workload]
    cfg_s, cfg_t = bench.make_cfgs(w, args.dropout)
    torch.manual_seed(1)
    student = magic_b200.GlocalTextPathCMTPreTraining(cfg_s).to(dev).train().set_compute_dtype(torch.bfloat16)
    teacher = None
    if cfg_t is not None:
        torch.manual_seed(0)
        teacher = magic_b200.GlocalTextPathCMTPreTraining(cfg_t).to(dev).eval().set_compute_dtype(torch.bfloat16)
    stepper = PretrainStepper(student, teacher, use_graphs=bool(args.graphs))
    ops.set_seed(dev, 1234)
    pools = {t: [batch_to_device(b, dev) for b in bench.make_pool(t, 2, w, 1234 + (0 if t == "mlm" else 500))]
             for t in ("mlm", "sap")}

    def pick(i):
        task = "mlm" if i % 2 == 0 else "sap"
        return task, pools[task][(i // 2) % 2]

    def step(i):
        task, b = pick(i)
        return stepper.step(task, b, next=pick(i + 1) if args.lookahead else None)

    for i in range(6):
        step(i)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(args.steps):
            step(i)
            if args.sync:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    rows = []
    for e in evs:
        tr = e.time_range
        rows.append((tr.start, tr.end - tr.start, e.name, getattr(e, "device_index", 0), getattr(e, "stream", -1)))
    rows.sort()
    t0 = rows[0][0] if rows else 0
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow(["start_us", "dur_us", "stream", "name"])
        for s, d, n, _, st in rows:
            wr.writerow([f"{s - t0:.3f}", f"{d:.3f}", st, n[:120]])
    print(f"{len(rows)} device events -> {args.out}")
    # split into steps at gaps > 50 us following a synchronize
    if not rows:
        return
    steps, cur = [], [rows[0]]
    for r in rows[1:]:
        prev_end = max(x[0] + x[1] for x in cur)
        if r[0] - prev_end > 40.0:
            steps.append(cur)
            cur = [r]
        else:
            cur.append(r)
    steps.append(cur)
    for k, st in enumerate(steps):
        a = st[0][0]
        b = max(x[0] + x[1] for x in st)
        # union busy time
        iv = sorted((x[0], x[0] + x[1]) for x in st)
        busy, ce = 0.0, a
        for s, e in iv:
            if e > ce:
                busy += e - max(s, ce)
                ce = e
        tot = sum(x[1] for x in st)
        print(f"segment {k}: {len(st)} kernels, wall {b - a:.1f} us, union-busy {busy:.1f} us, idle {b - a - busy:.1f} us, "
              f"sum of kernel time {tot:.1f} us")


if __name__ == "__main__":
    main()
