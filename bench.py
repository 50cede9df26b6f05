#!/usr/bin/env python
"""bench.py -- MAGIC pretraining hot path on B200 (contract: see the task statement / DESIGN.md section 6).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA kernels), one JSON line on rank 0
  python bench.py --impl reference ...                     the oracle port of the reference path on host cores

A "step" = one optimisation step (forward + loss + backward + [all-reduce] + clip + AdamW) of the MAGIC-S
student on one synthetic batch, tasks alternating MLM / SAP 1:1 (BASELINE.json configs[1]).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1]: MAGIC-S student pretraining step (MLM+SAP) bf16, batch 64
    "magic_s_pretrain_b64": dict(hidden=128, n_l=6, n_x=3, n_p=2, B=64, L=80, T_max=5, G_max=20, teacher=None),
    # BASELINE.json configs[2]: teacher h=768 (9/2/4) -> MAGIC-S distillation step, batch 64
    "magic_s_distill_t768_b64": dict(hidden=128, n_l=6, n_x=3, n_p=2, B=64, L=80, T_max=5, G_max=20,
                                     teacher=dict(hidden=768, n_l=9, n_x=4, n_p=2)),
    # MAGIC-L (h=768, 6/2/3) pretraining step without a teacher: the tensor-core-bound shape of the same path
    "magic_l_pretrain_b32": dict(hidden=768, n_l=6, n_x=3, n_p=2, B=32, L=80, T_max=5, G_max=20, teacher=None),
    # BASELINE.json configs[3] per GPU: MAGIC-L with ICoD teacher/student co-update (both models train), batch 32
    "magic_l_icod_b32": dict(hidden=768, n_l=6, n_x=3, n_p=2, B=32, L=80, T_max=5, G_max=20, co_update=True,
                             teacher=dict(hidden=768, n_l=9, n_x=4, n_p=2)),
    # BASELINE.json configs[4] per GPU: RxR-shape stress (160-token instr, 50-node graph, 12 steps), batch 128,
    # teacher h=768 -> MAGIC-S
    "rxr_stress_distill_b128": dict(hidden=128, n_l=6, n_x=3, n_p=2, B=128, L=160, T_max=12, G_max=50,
                                    teacher=dict(hidden=768, n_l=9, n_x=4, n_p=2)),
}
# forward GFLOP per sample, keyed (hidden, n_l, L): student / teacher shapes of the workloads above
FWD_GFLOP = {(128, 6, 80): 0.669, (768, 6, 80): 17.29, (768, 9, 80): 22.08, (128, 6, 160): 1.42, (768, 9, 160): 44.9}


def train_gflop_per_sample(w):
    """3 x forward for every model that trains + 1 x forward for a frozen teacher."""
    g = 3.0 * FWD_GFLOP.get((w["hidden"], w["n_l"], w["L"]), 0.0)
    t = w.get("teacher")
    if t:
        g += (3.0 if w.get("co_update") else 1.0) * FWD_GFLOP.get((t["hidden"], t["n_l"], w["L"]), 0.0)
    return g


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops"], tf_sus=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf=1590.0, tf_sus=1400.0, src="fallback")


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 2.0:  # nvidia-smi takes ~0.1 s to come up: be live first
                time.sleep(0.005)
            self.rows.clear()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=mx, reasons=sorted(reasons),
                    samples=len(sm))


def make_cfgs(w, dropout):
    from magic_b200.config import make_config
    t = w["teacher"]
    cfg_s = make_config(w["hidden"], w["n_l"], w["n_x"], w["n_p"], role="student",
                          teacher_hidden_size=t["hidden"] if t else None, hidden_dropout_prob=dropout,
                          attention_probs_dropout_prob=dropout)
    cfg_t = make_config(t["hidden"], t["n_l"], t["n_x"], t["n_p"], role="teacher", hidden_dropout_prob=0.0,
                          attention_probs_dropout_prob=0.0) if t else None
    return cfg_s, cfg_t


def make_pool(task, n, w, seed0):
    import magic_b200
    from magic_b200 import synth
    from magic_b200.graph_index import prepare_batch
    from magic_b200.graph_index import pad_batch
    out = []
    for i in range(n):
        b = synth.make_batch(task, w["B"], L=w["L"], T_max=w["T_max"], G_max=w["G_max"], seed=seed0 + i)
        out.append(prepare_batch(b))
    # identical shapes across the pool (one CUDA graph per task): pad to the pool maxima, rounded up
    rcap = max(b["traj_view_img_fts"].shape[0] for b in out)
    rcap = (rcap + 7) // 8 * 8
    mcap = None
    if task == "mlm":
        mcap = max(b[magic_b200.INDEX_KEY]["mlm_rows"].numel() for b in out)
        mcap = (mcap + 63) // 64 * 64
    K = magic_b200.INDEX_KEY
    ecap = (max(b[K]["entries"].numel() for b in out) + 255) // 256 * 256
    scap = (max(b[K]["src_ids"].numel() for b in out) + 255) // 256 * 256
    return [pad_batch(b, rcap, mcap, ecap, scap) for b in out]


def host_pin(batch):
    import magic_b200
    out = {}
    for k, v in batch.items():
        if torch.is_tensor(v):
            out[k] = v.pin_memory()
        elif k == magic_b200.INDEX_KEY:
            out[k] = {kk: (vv.pin_memory() if torch.is_tensor(vv) else vv) for kk, vv in v.items()}
        else:
            out[k] = v
    return out


def nbytes(batch):
    import magic_b200
    n = 0
    for k, v in batch.items():
        if k == magic_b200.FLAT_KEY:
            continue
        if torch.is_tensor(v):
            n += v.numel() * v.element_size()
        elif k == magic_b200.INDEX_KEY:
            n += sum(vv.numel() * vv.element_size() for vv in v.values() if torch.is_tensor(vv))
    return n


# ---------------------------------------------------------------------------------------------------
# roofline bookkeeping for the instrumented pass
# ---------------------------------------------------------------------------------------------------
def family_cost(name, a):
    """-> (flops, bytes) of one C-ABI call from its argument tuple (algorithmic)."""
    esz = lambda dtc: 2 if dtc == 1 else 4
    if name == "magic_gemm":
        M, N, K = a[11], a[12], a[13]
        return 2.0 * M * N * K, M * K * esz(a[1]) + N * K * esz(a[5]) + M * N * esz(a[9])
    if name == "magic_gemm_wgrad":
        M, N, K = a[9], a[10], a[11]
        return 2.0 * M * N * K, M * N * esz(a[1]) + M * K * esz(a[4]) + N * K * 4
    if name in ("magic_attn_fwd", "magic_attn_bwd"):
        if name == "magic_attn_fwd":
            B, H, Lq, Lk, dtc = a[11], a[12], a[13], a[14], a[20]
            mul = 1
        else:
            B, H, Lq, Lk, dtc = a[19], a[20], a[21], a[22], a[28]
            mul = 2.5
        return 4.0 * B * H * Lq * Lk * 64 * mul, (2 * B * Lq + 2 * B * Lk) * H * 64 * esz(dtc) * (2 if mul > 1 else 1)
    if name in ("magic_ln_fwd", "magic_ln_bwd"):
        M, h, dtc = (a[6], a[7], a[9]) if name == "magic_ln_fwd" else (a[9], a[10], a[11])
        return 0.0, M * h * esz(dtc) * (3 if name == "magic_ln_fwd" else 5)
    if name in ("magic_makd_mse_fwd", "magic_makd_mse_bwd"):
        by = 0.0
        for i in range(a[1]):
            sg = a[0][i]
            n = sg.rows * sg.inner
            by += n * (esz(sg.s_dt) + esz(sg.t_dt)) + (n * esz(sg.s_dt) if name.endswith("bwd") else 0)
        return 0.0, by
    if name in ("magic_makd_kl_fwd", "magic_makd_kl_bwd"):
        R, C = a[2], a[3] if name.endswith("fwd") else a[4]
        if name.endswith("bwd"):
            R, C, dtc = a[3], a[4], a[-2]
            return 0.0, R * C * esz(dtc) * 3
        dtc = a[-2]
        return 0.0, R * C * esz(dtc) * 2 * 2   # two passes over student + teacher rows
    if name in ("magic_ce_fwd", "magic_ce_bwd"):
        R, C, dtc = (a[4], a[5], a[8]) if name.endswith("fwd") else (a[5], a[6], a[9])
        return 0.0, R * C * esz(dtc) * (1 if name.endswith("fwd") else 2)
    if name == "magic_colsum":
        return 0.0, a[2] * a[3] * esz(a[5])
    if name == "magic_adamw":
        return 0.0, a[5] * (28 + (2 if a[4] else 0))
    if name == "magic_sumsq":
        return 0.0, a[1] * 4
    return 0.0, 0.0


OUTLIERS = {}
WORKLOAD_NAME = None


def summarise_profile(prof, n_steps, pk):
    fam = {}
    for name, recs in prof.items():
        if name == "magic_delay":
            continue
        per_call = [e0.elapsed_time(e1) for e0, e1, _ in recs]
        # one-off stalls inside a bracket (seen: a single 39 ms gap in one attention-backward call of the eager
        # distillation pass, 1000x its median) are not kernel time: a call above max(20 x median, 1 ms) counts as
        # the family's median and is reported in `profile_outliers`
        med = sorted(per_call)[len(per_call) // 2]
        lim = max(20.0 * med, 1.0)
        n_out = sum(1 for x in per_call if x > lim)
        if n_out:
            OUTLIERS[name] = OUTLIERS.get(name, 0) + n_out
            per_call = [med if x > lim else x for x in per_call]
        ms = sum(per_call)
        if os.environ.get("BENCH_DEBUG_PROFILE"):
            srt = sorted(per_call)
            sys.stderr.write(f"[profile] {name}: n={len(srt)} sum={ms:.3f} ms median={srt[len(srt) // 2] * 1e3:.1f} us "
                             f"max={srt[-1] * 1e3:.1f} us top5={[round(x * 1e3, 1) for x in srt[-5:]]}\n")
        fl = by = 0.0
        for _, _, a in recs:
            f, b = family_cost(name, a)
            fl += f
            by += b
        key = "magic_gemm" if name == "magic_gemm_wgrad" else name  # one kernel (gemm_tc_kernel), one family
        d = fam.setdefault(key, dict(ms_per_step=0.0, calls_per_step=0.0, flops=0.0, bytes=0.0))
        d["ms_per_step"] += ms / n_steps
        d["calls_per_step"] += len(recs) / n_steps
        d["flops"] += fl / n_steps
        d["bytes"] += by / n_steps
    tot = sum(v["ms_per_step"] for v in fam.values()) or 1.0
    for v in fam.values():
        v["share"] = v["ms_per_step"] / tot
    top = max(fam.items(), key=lambda kv: kv[1]["ms_per_step"])
    name, v = top
    if v["flops"] > 0 and name == "magic_gemm":
        ach = v["flops"] / (v["ms_per_step"] * 1e-3) / 1e12
        roof = dict(kernel=name, bound="tensor", achieved=ach, peak=pk["tf_sus"], unit="TFLOP/s",
                    frac=ach / pk["tf_sus"], traffic=None, peak_source=pk["src"] + " (sustained)")
    else:
        ach = v["bytes"] / (v["ms_per_step"] * 1e-3) / 1e9
        roof = dict(kernel=name, bound="hbm", achieved=ach, peak=pk["hbm"], unit="GB/s", frac=ach / pk["hbm"],
                    traffic=None, peak_source=pk["src"])
    # DRAM traffic per launch of the dominant kernel from the committed `ncu --set full` capture of this workload
    # (profiles/traffic.json, written by scripts/summarize_ncu.py traffic); null when no capture covers it
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = tr.get(WORKLOAD_NAME, {}).get(name)
        if ent:
            roof["traffic"] = ent["dram_bytes_per_launch"]
            roof["traffic_source"] = ent["source"]
            roof["algorithmic_bytes_per_launch"] = v["bytes"] / max(v["calls_per_step"], 1)
    except Exception:
        pass
    roof["avg_launch_us"] = v["ms_per_step"] * 1e3 / max(v["calls_per_step"], 1)
    roof["share_of_step_kernel_time"] = v["share"]
    fams = {k: dict(ms_per_step=round(x["ms_per_step"], 4), calls=round(x["calls_per_step"], 1),
                    share=round(x["share"], 3),
                    tflops=round(x["flops"] / (x["ms_per_step"] * 1e-3) / 1e12, 2) if x["flops"] else None,
                    gbs=round(x["bytes"] / (x["ms_per_step"] * 1e-3) / 1e9, 1) if x["bytes"] else None)
            for k, x in sorted(fam.items(), key=lambda kv: -kv[1]["ms_per_step"])[:8]}
    return roof, fams


# ---------------------------------------------------------------------------------------------------
# CPU arms
# ---------------------------------------------------------------------------------------------------
def cpu_step_rate(w, seconds_budget, sample_B, dropout=0.0, train=True):
    """Oracle port (fp32 PyTorch) of the same step on the host cores: student fwd+bwd+AdamW, MLM/SAP 1:1."""
    from oracle import magic_oracle as O
    from magic_b200 import synth
    torch.set_num_threads(os.cpu_count())
    cfg_s, _ = make_cfgs(w, dropout)
    torch.manual_seed(1)
    model = O.GlocalTextPathCMTPreTraining(cfg_s).train()
    opt = torch.optim.AdamW(model.parameters(), lr=5e-5, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.01)
    batches = [(t, synth.make_batch(t, sample_B, L=w["L"], T_max=w["T_max"], G_max=w["G_max"], seed=9 + i))
               for i, t in enumerate(("mlm", "sap"))]
    times, n, t_start = [], 0, time.time()
    while True:
        task, b = batches[n % 2]
        t0 = time.time()
        opt.zero_grad()
        loss = model(b, task, True)["loss"].mean()
        if train:
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0)
            opt.step()
        times.append(time.time() - t0)
        n += 1
        if n >= 4 and (time.time() - t_start > seconds_budget or n >= 40):
            break
    t = statistics.median(times[2:]) if len(times) > 3 else statistics.median(times)
    return sample_B / t, n, t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    sample_B = 8
    times = []
    from oracle import magic_oracle as O
    from magic_b200 import synth
    torch.set_num_threads(os.cpu_count())
    cfg_s, _ = make_cfgs(w, args.dropout)
    torch.manual_seed(1)
    model = O.GlocalTextPathCMTPreTraining(cfg_s).train()
    opt = torch.optim.AdamW(model.parameters(), lr=5e-5, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.01)
    batches = [(t, synth.make_batch(t, sample_B, L=w["L"], T_max=w["T_max"], G_max=w["G_max"], seed=9 + i))
               for i, t in enumerate(("mlm", "sap"))]
    for i in range(args.warmup + args.steps):
        task, b = batches[i % 2]
        t0 = time.time()
        opt.zero_grad()
        loss = model(b, task, True)["loss"].mean()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0)
        opt.step()
        if i >= args.warmup:
            times.append(time.time() - t0)
    tot = sum(times)
    v = sample_B * len(times) / tot
    sample = f"fp32 PyTorch oracle port (reference model files absent upstream), student step fwd+bwd+clip+AdamW, " \
             f"batch {sample_B} per step (bounded sample of the batch-{w['B']} workload), MLM/SAP 1:1"
    emit(OUT_FD, (dict(
        impl="reference", metric="pretrain samples/s (MLM+SAP step)", value=v, unit="samples/s", n_gpus=args.gpus,
        steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * tot / len(times), higher_is_better=True,
        scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
        config=dict(workload=args.workload, hidden=w["hidden"], layers=f"{w['n_l']}/{w['n_p']}/{w['n_x']}",
                    batch_per_step=sample_B, seq_len=w["L"], graph_nodes=w["G_max"]),
        cpu_baseline=dict(value=v, unit="samples/s", cores=os.cpu_count(), kind="port", sample=sample),
        e2e=dict(value=v, unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))))


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import magic_b200
    from magic_b200 import _lib, ops
    from magic_b200.graph_index import flatten_batch
    from magic_b200.train_step import PretrainStepper

    from magic_b200.parallel import init_distributed
    rank, world, local = init_distributed()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    w = WORKLOADS[args.workload]
    global WORKLOAD_NAME
    WORKLOAD_NAME = args.workload
    pk = peaks()
    cfg_s, cfg_t = make_cfgs(w, args.dropout)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    torch.manual_seed(1)
    student = magic_b200.GlocalTextPathCMTPreTraining(cfg_s).to(dev).train().set_compute_dtype(dtype)
    teacher = None
    if cfg_t is not None:
        torch.manual_seed(0)
        teacher = magic_b200.GlocalTextPathCMTPreTraining(cfg_t).to(dev).set_compute_dtype(dtype)
        teacher = teacher.train() if w.get("co_update") else teacher.eval()
    stepper = PretrainStepper(student, teacher, use_graphs=bool(args.graphs), co_update=bool(w.get("co_update")),
                              side_stream=bool(args.side_stream), branch_streams=bool(args.branch_streams))
    ops.set_seed(dev, 1234 + rank)

    pool_n = args.pool
    pools = {t: make_pool(t, pool_n, w, 1234 + rank * 1000 + (0 if t == "mlm" else 500)) for t in ("mlm", "sap")}
    # flat batches: every tensor of a batch is a view into one buffer, so staging a batch is ONE copy
    dev_pools = {t: [flatten_batch(b, device=dev) for b in bs] for t, bs in pools.items()}
    pin_pools = {t: [flatten_batch(b, pin=True) for b in bs] for t, bs in pools.items()}
    in_bytes = sum(nbytes(b) for bs in pools.values() for b in bs) / (2 * pool_n)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_loss = [torch.zeros(1).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]

    def pick(i):
        return ("mlm" if i % 2 == 0 else "sap"), (i // 2) % pool_n

    def run_steps(n, first, from_host, delay_cycles=0):
        out = None
        nxt = None
        if from_host:
            task, j = pick(first)
            nxt = stepper.prefetch(task, pin_pools[task][j])
        for i in range(first, first + n):
            if delay_cycles:
                # keep the GPU busy while the host queues this step's launches, so per-call CUDA events bracket
                # back-to-back device execution rather than host launch latency
                _lib.COUNTERS["launches"] -= 1
                _lib.call("magic_delay", int(delay_cycles), _lib.stream())
            task, j = pick(i)
            if from_host:
                # end to end: every step's inputs start in pinned host memory; the copy of step i+1 is issued on the
                # copy stream before step i's loss is read back, as the reference's PrefetchLoader does
                b = nxt
                if i + 1 < first + n:
                    t2, j2 = pick(i + 1)
                    nxt = stepper.prefetch(t2, pin_pools[t2][j2])
            else:
                b = dev_pools[task][j]
            out = stepper.step(task, b)
            if from_host:
                # device -> host read of EVERY step's loss, one step late (async copy into pinned memory + event),
                # so the host queues step i+1 while step i runs instead of draining the GPU each step
                k = i % 2
                host_loss[k].copy_(out[0:1], non_blocking=True)
                loss_ev[k].record()
                if i > first:
                    loss_ev[1 - k].synchronize()
                    _ = host_loss[1 - k].item()
        if from_host:
            loss_ev[(first + n - 1) % 2].synchronize()
            _ = host_loss[(first + n - 1) % 2].item()
        return out

    # warm-up (also builds the CUDA graphs, one per task/shape)
    run_steps(max(args.warmup, 3), 0, False)
    sync_all()
    clk = ClockSampler(local)
    if rank == 0:
        clk.start()
    c0 = dict(_lib.COUNTERS)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    run_steps(args.steps, 100, False)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = _lib.COUNTERS["launches"] - c0["launches"]
    if args.timed_only:
        clk.stop() if rank == 0 else None
        sys.stderr.write(f"timed-only: {ms / args.steps:.3f} ms/step, {launches} launches\n")
        if world > 1:
            dist.destroy_process_group()
        return
    # end-to-end: host (pinned) buffers -> H2D -> step -> D2H loss, through the public stepper API
    run_steps(2, 0, True)
    sync_all()
    t0 = time.perf_counter()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    run_steps(args.steps, 100, True)
    f1.record()
    sync_all()
    ms_e2e = max(f0.elapsed_time(f1), (time.perf_counter() - t0) * 1e3)
    clocks = clk.stop() if rank == 0 else None  # sampled (20 ms period) across BOTH timed regions: device-resident and e2e
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    B = w["B"]
    value = world * B * args.steps / (ms * 1e-3)
    e2e_v = world * B * args.steps / (ms_e2e * 1e-3)

    roof, fams = None, None
    if rank == 0:
        # instrumented pass (eager, per-call CUDA events on the launch stream) for the roofline numbers
        # (single stream, eager; a delay kernel every 32 calls keeps the host ahead of the GPU -- see _lib.profile_start)
        g = stepper.use_graphs
        stepper.use_graphs = False
        # rank 0 runs this pass ALONE: it must not issue the gradient all-reduce (the other ranks are not in it)
        stepper.allreduce = stepper.t_allreduce = None
        ops.enable_side_stream(False)
        ops.enable_branch_streams(False)
        run_steps(2, 0, False)
        torch.cuda.synchronize()
        _lib.profile_start(delay_every=32, delay_cycles=2e6)
        nprof = 4
        run_steps(nprof, 100, False)
        torch.cuda.synchronize()
        roof, fams = summarise_profile(_lib.profile_stop(), nprof, pk)
        stepper.use_graphs = g
        ops.enable_side_stream(bool(args.side_stream))
        ops.enable_branch_streams(bool(args.branch_streams))
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, n, tstep = cpu_step_rate(w, 12.0, 8, args.dropout)
        cpu = dict(value=v, unit="samples/s", cores=os.cpu_count(), kind="port",
                   sample=f"fp32 PyTorch oracle port, student step fwd+bwd+clip+AdamW at batch 8 (bounded sample of the "
                          f"batch-{B} workload), MLM/SAP 1:1, {n} steps, median {tstep * 1e3:.0f} ms/step")
    if rank == 0:
        train_gflop = train_gflop_per_sample(w)
        kind = "MLM+SAP step" if w["teacher"] is None else ("MLM+SAP+distill step, ICoD co-update" if w.get("co_update")
                                                            else "MLM+SAP+distill step")
        line = dict(
            metric=f"pretrain samples/s ({kind})", value=value, unit="samples/s", n_gpus=world, steps=args.steps,
            warmup=max(args.warmup, 3), ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak",
            vs_baseline=None, dtype=args.dtype, data="synthetic",
            config=dict(workload=args.workload, hidden=w["hidden"], layers=f"{w['n_l']}/{w['n_p']}/{w['n_x']}",
                        batch_per_gpu=B, global_batch=B * world, seq_len=w["L"], views=36, graph_nodes=w["G_max"],
                        traj_steps_max=w["T_max"], tasks="mlm:sap 1:1", dropout=args.dropout,
                        teacher=("h%d %d/%d/%d%s" % (w["teacher"]["hidden"], w["teacher"]["n_l"], w["teacher"]["n_p"],
                                                     w["teacher"]["n_x"], " (trained, ICoD)" if w.get("co_update")
                                                     else " (frozen)")) if w["teacher"] else None,
                        optimizer="fused AdamW + clip 5.0", cuda_graphs=bool(stepper.use_graphs),
                        parallelism=f"dp{world}",
                        l2="inputs cycle through a pool of %d batches/task (~%.0f MB) > 126 MB L2" % (
                            pool_n, 2 * pool_n * in_bytes / 1e6),
                        model_tflops=value * train_gflop / 1e3),
            roofline=roof, kernel_families=fams, profile_outliers=OUTLIERS or None, cpu_baseline=cpu,
            e2e=dict(value=e2e_v, unit="samples/s", h2d_bytes_per_step=int(in_bytes), d2h_bytes_per_step=4,
                     ms_per_step=ms_e2e / args.steps),
            gpu_launches=int(launches), clocks=clocks)
        emit(OUT_FD, line)
    if world > 1:
        dist.destroy_process_group()


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to fd 1
    when NCCL_DEBUG is set), so fd 1 is pointed at stderr for the whole run and the result line goes to the saved
    descriptor."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return saved


def emit(saved_fd, obj):
    os.write(saved_fd, (json.dumps(obj) + "\n").encode())


OUT_FD = None


def main():
    global OUT_FD
    OUT_FD = claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="magic_s_pretrain_b64", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--graphs", type=int, default=1)
    ap.add_argument("--pool", type=int, default=8)
    ap.add_argument("--side-stream", type=int, default=1)
    ap.add_argument("--branch-streams", type=int, default=1)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--timed-only", action="store_true",
                    help="profiling aid: run warm-up + the timed region only (no e2e / roofline / CPU passes, no JSON)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
