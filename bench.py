#!/usr/bin/env python
"""bench.py -- MAGIC pretraining hot path on B200 (contract: see the task statement / DESIGN.md section 6).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA kernels), one JSON line on rank 0
  python bench.py --impl reference ...                     the oracle port of the reference path on host cores

  python bench.py --impl torch_gpu ...                     the same oracle port with stock torch ops on cuda:0 (bf16 autocast)

A "step" = one optimisation step of the distillation hot path on one synthetic batch: frozen teacher (h = 768, 9/2/4)
forward, MAGIC-S student forward, MAKD losses, backward, [gradient exchange], clip + AdamW; tasks alternate MLM / SAP
1:1 (BASELINE.json configs[2], the configuration the metric "MLM+SAP+distill step" names).  The other configs ride
along in the `workloads` sub-dict of the same JSON line.
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
# the headline: BASELINE.json's metric is "pretrain samples/s (MLM+SAP+distill step)" = configs[2]
DEFAULT_WORKLOAD = "magic_s_distill_t768_b64"
# BASELINE.json configs[1], [2], [3], [4] (in that order): carried in the JSON line's `workloads` sub-dict
SUB_WORKLOADS = ("magic_s_pretrain_b64", "magic_s_distill_t768_b64", "magic_l_icod_b32", "rxr_stress_distill_b128")
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


def make_pool(task, n, w, seed0, store=None):
    """`store`: a CPU copy of the panorama feature database; the batches then carry panorama rows + view orders instead
    of the 36 x 768 features per step (featurizer.py) and the model gathers the features on the device."""
    import magic_b200
    from magic_b200 import synth
    from magic_b200.featurizer import compact_batch
    from magic_b200.graph_index import prepare_batch
    from magic_b200.graph_index import pad_batch
    out = []
    for i in range(n):
        b = synth.make_batch(task, w["B"], L=w["L"], T_max=w["T_max"], G_max=w["G_max"], seed=seed0 + i, store=store)
        if store is not None:
            b = compact_batch(b)
        out.append(prepare_batch(b))
    # identical shapes across the pool (one CUDA graph per task): pad to the pool maxima, rounded up
    rcap = max(b["traj_vp_view_lens"].shape[0] for b in out)
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
    if name in ("magic_attn_fwd", "magic_attn_bwd", "magic_attn_bwd_part"):
        if name == "magic_attn_fwd":
            B, H, Lq, Lk, dtc = a[11], a[12], a[13], a[14], a[20]
            mul = 1
        elif name == "magic_attn_bwd":
            B, H, Lq, Lk, dtc = a[19], a[20], a[21], a[22], a[28]
            mul = 2.5
        else:  # one half (query-major or key-major) of the split backward
            B, H, Lq, Lk, dtc = a[17], a[18], a[19], a[20], a[26]
            mul = 1.25
        # bytes: Q, O (+dO, dQ) and K, V (+dK, dV) once each; a backward half touches half of the backward's tensors
        return 4.0 * B * H * Lq * Lk * 64 * mul, (2 * B * Lq + 2 * B * Lk) * H * 64 * esz(dtc) * (2 if mul > 2 else 1)
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
        if name.endswith("bwd"):
            R, C, dtc = a[3], a[4], a[-2]
            return 0.0, R * C * esz(dtc) * 3   # reads student + teacher rows, writes dS
        R, C, dtc = a[2], a[3], a[11]
        return 0.0, R * C * esz(dtc) * 2       # single pass: every student / teacher logit is read once
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


WORKLOAD_NAME = None


# ---------------------------------------------------------------------------------------------------
# baselines: the oracle port on the host cores (reference arm / cpu_baseline) and on the GPU (gpu_baseline)
# ---------------------------------------------------------------------------------------------------
def oracle_step_fn(w, device, dropout, autocast=False, fused_opt=False):
    """-> (step(i) -> None, batch size): one optimisation step of workload `w` written with stock PyTorch ops only
    (the oracle port of the missing reference model + the reference's kd_loss arithmetic): frozen-teacher forward,
    student forward, MAKD losses, alpha mix, backward, clip 5.0, AdamW; ICoD co-update trains both models.
    MLM / SAP alternate 1:1."""
    from oracle import magic_oracle as O
    from magic_b200 import synth
    cfg_s, cfg_t = make_cfgs(w, dropout)
    torch.manual_seed(1)
    student = O.GlocalTextPathCMTPreTraining(cfg_s).to(device).train()
    teacher = None
    co = bool(w.get("co_update"))
    if cfg_t is not None:
        torch.manual_seed(0)
        teacher = O.GlocalTextPathCMTPreTraining(cfg_t).to(device)
        teacher = teacher.train() if co else teacher.eval()
    kw = dict(lr=5e-5, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.01)
    if fused_opt:
        kw["fused"] = True
    opts = [torch.optim.AdamW(student.parameters(), **kw)]
    if co:
        opts.append(torch.optim.AdamW(teacher.parameters(), **kw))
    B = w["B"]
    batches = []
    for i, t in enumerate(("mlm", "sap")):
        b = synth.make_batch(t, B, L=w["L"], T_max=w["T_max"], G_max=w["G_max"], seed=9 + i)
        batches.append((t, {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in b.items()}))
    gen = torch.Generator().manual_seed(5)

    def step(i):
        task, b = batches[i % 2]
        for o in opts:
            o.zero_grad(set_to_none=True)
        with torch.autocast(device_type="cuda", dtype=torch.bfloat16, enabled=autocast):
            if teacher is None:
                loss = student(b, task, True)["loss"].float().mean()
            else:
                rw = O.mkrw_weights(gen, 4.0, device)
                if co:
                    tot_s, tot_t = O.icod_step_loss(student, teacher, b, task, rw, rw)[:2]
                    loss = tot_s.float() + tot_t.float()
                else:
                    loss = O.distill_step_loss(student, teacher, b, task, rw)[0].float()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(student.parameters(), 5.0)
        if co:
            torch.nn.utils.clip_grad_norm_(teacher.parameters(), 5.0)
        for o in opts:
            o.step()

    return step, B


def kind_of(w):
    return "MLM+SAP step" if w["teacher"] is None else ("MLM+SAP+distill step, ICoD co-update" if w.get("co_update")
                                                        else "MLM+SAP+distill step")


def cpu_step_rate(w, seconds_budget, dropout=0.0):
    """Bounded sample of the workload on the host cores (same batch size, same models): >= 2 steps (one MLM, one SAP),
    more while the budget lasts."""
    torch.set_num_threads(os.cpu_count())
    step, B = oracle_step_fn(w, "cpu", dropout)
    times, t_start = [], time.time()
    while True:
        t0 = time.time()
        step(len(times))
        times.append(time.time() - t0)
        if len(times) >= 2 and (time.time() - t_start > seconds_budget or len(times) >= 40):
            break
    use = times[1:] if len(times) > 2 else times  # the first step pays the allocator / thread-pool start-up
    t = sum(use) / len(use)
    return B / t, len(times), t


def run_reference(args):
    """The reference arm: the oracle port (the reference's model files are absent upstream, readme.md:75; its KD loss
    arithmetic is the reference's own) running THE SAME workload -- same models, same batch size, same step --
    in fp32 on all host cores.  Rank 0 alone runs it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    torch.set_num_threads(os.cpu_count())
    step, B = oracle_step_fn(w, "cpu", args.dropout)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.time()
        step(i)
        if i >= args.warmup:
            times.append(time.time() - t0)
    tot = sum(times)
    v = B * len(times) / tot
    t = w["teacher"]
    sample = f"fp32 PyTorch oracle port (reference model files absent upstream) of the same step at the same batch " \
             f"size {B}: " + ("frozen teacher forward + " if t and not w.get("co_update") else "") + \
             "student forward/backward" + (" + teacher forward/backward (ICoD)" if w.get("co_update") else "") + \
             (" + MAKD losses" if t else "") + " + clip + AdamW, MLM/SAP 1:1"
    emit(OUT_FD, (dict(
        impl="reference", metric=f"pretrain samples/s ({kind_of(w)})", value=v, unit="samples/s", n_gpus=args.gpus,
        steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * tot / len(times), higher_is_better=True,
        scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
        config=workload_config(args.workload, w, 1, args.dropout, cuda_graphs=False),
        cpu_baseline=dict(value=v, unit="samples/s", cores=os.cpu_count(), kind="port", sample=sample),
        e2e=dict(value=v, unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))))


def torch_gpu_rate(w, dropout, steps=6, warmup=2, compiled=False):
    """Stock-PyTorch-on-GPU baseline (SURVEY.md 8d "the number our sm_100a kernels must beat"): the same oracle port
    on cuda:0 under bf16 autocast (cuBLASLt GEMMs, ATen softmax / LayerNorm, fused torch AdamW), eager."""
    step, B = oracle_step_fn(w, "cuda", dropout, autocast=True, fused_opt=True)
    if compiled:
        step = torch.compile(step, dynamic=False)
    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return B / (ms * 1e-3), ms


def run_torch_gpu(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    torch.cuda.set_device(0)
    v, ms = torch_gpu_rate(w, args.dropout, max(args.steps, 2), max(args.warmup, 2), compiled=bool(args.compile))
    emit(OUT_FD, dict(impl="torch_gpu", metric=f"pretrain samples/s ({kind_of(w)})", value=v, unit="samples/s",
                      n_gpus=1, steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=True,
                      scaling="weak", vs_baseline=None, dtype="bf16 autocast", data="synthetic",
                      config=workload_config(args.workload, w, 1, args.dropout, cuda_graphs=False),
                      kind="oracle port, stock torch ops on cuda" + (", torch.compile" if args.compile else ", eager")))


def nav_latency(dtype=torch.bfloat16, steps=12):
    """Fine-tune / inference path (the "MAGIC-S real-time" claim): one navigation decision -- panorama
    mode + GraphMap update + collators + navigation mode + read-back of the action -- of nav.VLNBert (MAGIC-S, h = 128)
    over online GraphMaps on synthetic worlds, batch 1 and 8, 80-token instructions, graphs growing to ~35 nodes.
    Eager (`vln_bert(mode, batch)`, the reference's call) and graph-replayed (nav.NavStepper); host python included in
    the end-to-end figure, CUDA events around the navigation mode for the device figure."""
    import numpy as np
    from magic_b200 import nav, nav_synth
    from magic_b200.config import make_config

    def rollout(model, stepper, B, replay):
        rng = np.random.RandomState(5)
        worlds = [nav_synth.NavWorld(n=40, seed=60 + b) for b in range(B)]
        obs = [w.observe(int(rng.randint(0, 40)), instr=nav_synth.make_instr(rng, 80)) for w in worlds]
        gmaps = [nav.GraphMap(ob["viewpoint"]) for ob in obs]
        for gm, ob in zip(gmaps, obs):
            gm.update_graph(ob)
        lang = nav.language_inputs(obs, "cuda")
        txt, _ = model("language", lang)
        last, times, nodes = None, [], 0
        for t in range(steps + 3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for gm, ob in zip(gmaps, obs):
                gm.node_step_ids[ob["viewpoint"]] = t + 1
            pin = nav.panorama_inputs(obs, "cuda")
            pe, pm, pf, _ = stepper.panorama(pin) if replay else model("panorama", pin)
            for i, (gm, ob) in enumerate(zip(gmaps, obs)):
                gm.update_node_embed(ob["viewpoint"], pf[i], rewrite=True)
                for j, c in enumerate(pin["cand_vpids"][i]):
                    if not gm.graph.visited(c):
                        gm.update_node_embed(c, pe[i, j])
            nin = nav.nav_gmap_inputs(obs, gmaps, last)
            nin.update(nav.nav_vp_inputs_mem(obs, gmaps, pe, pin["cand_vpids"], pin["view_lens"], pin["nav_types"], last))
            nin.update(txt_embeds=txt, txt_masks=lang["txt_masks"], txt_lens=lang["txt_lens"])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            o = stepper.navigation(nin) if replay else model("navigation", nin)
            e1.record()
            act = o["fused_logits"].argmax(1).tolist()  # device -> host read of the decision
            torch.cuda.synchronize()
            if t >= 3:
                times.append(((time.perf_counter() - t0) * 1e3, e0.elapsed_time(e1)))
            last = o["cls_embeds"].clone()
            nodes = int(nin["gmap_masks"].shape[1])
            nxt = []
            for i, (w, ob) in enumerate(zip(worlds, obs)):
                vp = nin["gmap_vpids"][i][act[i]]
                nb = np.nonzero(w.adj[w.index(ob["viewpoint"])])[0]
                j = w.index(vp) if vp is not None else int(nb[t % len(nb)])
                nxt.append(w.observe(j, heading=0.3 * t, instr=ob["instr_encoding"]))
            obs = nxt
            for gm, ob in zip(gmaps, obs):
                gm.update_graph(ob)
        wall = sorted(x[0] for x in times)[len(times) // 2]
        dev = sorted(x[1] for x in times)[len(times) // 2]
        return wall, dev, nodes

    out = {}
    for B in (1, 8):
        cfg = make_config(128, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, pretrain_tasks=("sap",))
        torch.manual_seed(0)
        model = nav.VLNBert(cfg).to("cuda").eval().set_compute_dtype(dtype)
        model.want_attn = False
        stepper = nav.NavStepper(model, B, G=64, Lt=80)
        with torch.no_grad():
            ew, ed, nodes = rollout(model, stepper, B, False)
            gw, gd, _ = rollout(model, stepper, B, True)
        out[f"batch{B}"] = dict(eager=dict(ms_per_decision_end_to_end=round(ew, 3), ms_navigation_mode_device=round(ed, 3)),
                                graph_replay=dict(ms_per_decision_end_to_end=round(gw, 3),
                                                  ms_navigation_mode_device=round(gd, 3)),
                                decisions_per_s=round(B * 1e3 / gw, 1), graph_nodes=nodes)
    out["what"] = ("nav.VLNBert MAGIC-S bf16: panorama mode + GraphMap update + collators + navigation mode + argmax read-back "
                   "per decision (median of %d), host python included in the end-to-end figure" % steps)
    return out


def featurizer_rate(B=64, Tmax=5, calls=20):
    """GPU batch featuriser, graph half: device time of one magic_featurize_graph call (the graph-
    derived tensors and index tables of a B-sample batch) vs the host loops it replaces (graph_index.build_index on the
    collated batch; the reference's per-sample dataset.py code is slower still)."""
    import numpy as np
    from magic_b200 import nav_synth, synth
    from magic_b200.featurizer import GraphFeaturizer, GraphWorld
    from magic_b200.graph_index import build_index
    w = nav_synth.NavWorld(n=256, seed=9)
    world = GraphWorld(*nav_synth.world_tables(w), device="cuda")
    rng = np.random.RandomState(3)
    paths = []
    for b in range(B):
        p = [int(rng.randint(0, w.n))]
        for _ in range(int(rng.randint(1, Tmax))):
            nb = np.nonzero(w.adj[p[-1]])[0][:8]
            p.append(int(nb[rng.randint(len(nb))]))
        paths.append(p)
    feat = GraphFeaturizer(world, B, Tmax=Tmax, G=64)
    heads = [0.1 * b for b in range(B)]
    nxt = [-1] * B
    feat(paths, heads, nxt)
    feat.check()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(calls):
        feat(paths, heads, nxt)
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / calls
    dev = e0.elapsed_time(e1) / calls
    hb = synth.make_batch("sap", B, seed=1, T_max=Tmax)
    t0 = time.perf_counter()
    for _ in range(3):
        build_index(hb)
    host = (time.perf_counter() - t0) * 1e3 / 3
    return dict(batch=B, ms_per_batch_device=round(dev, 3), ms_per_batch_wall=round(wall, 3),
                host_build_index_ms=round(host, 3), h2d_bytes_per_batch=int(B * (Tmax + 5) * 4),
                what="magic_featurize_graph (2 kernels) on a 256-viewpoint synthetic world vs graph_index.build_index "
                     "(python loops over the collated id strings) for the same batch size")


def workload_config(name, w, world, dropout, cuda_graphs=True, pool_n=None, in_bytes=None):
    B = w["B"]
    cfg = dict(workload=name, hidden=w["hidden"], layers=f"{w['n_l']}/{w['n_p']}/{w['n_x']}",
               batch_per_gpu=B, global_batch=B * world, seq_len=w["L"], views=36, graph_nodes=w["G_max"],
               traj_steps_max=w["T_max"], tasks="mlm:sap 1:1", dropout=dropout,
               teacher=("h%d %d/%d/%d%s" % (w["teacher"]["hidden"], w["teacher"]["n_l"], w["teacher"]["n_p"],
                                            w["teacher"]["n_x"], " (trained, ICoD)" if w.get("co_update")
                                            else " (frozen)")) if w["teacher"] else None,
               optimizer="fused AdamW + clip 5.0", cuda_graphs=bool(cuda_graphs), parallelism=f"dp{world}")
    if pool_n is not None:
        cfg["l2"] = ("the step's working set (weights + activations + logits, > 1 GB at h = 768) exceeds the 126 MB L2; "
                     "inputs cycle through a pool of %d batches/task" % pool_n)
    return cfg


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
class Runner:
    """One workload on this rank: models, stepper, input pools, and the step loops bench.py times."""

    def __init__(self, name, args, rank, world, dev):
        import magic_b200
        from magic_b200 import ops
        from magic_b200.graph_index import flatten_batch
        from magic_b200.train_step import PretrainStepper
        self.name, self.args, self.rank, self.world, self.dev = name, args, rank, world, dev
        w = self.w = WORKLOADS[name]
        cfg_s, cfg_t = make_cfgs(w, args.dropout)
        dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
        torch.manual_seed(1)
        student = magic_b200.GlocalTextPathCMTPreTraining(cfg_s).to(dev).train().set_compute_dtype(dtype)
        teacher = None
        if cfg_t is not None:
            torch.manual_seed(0)
            teacher = magic_b200.GlocalTextPathCMTPreTraining(cfg_t).to(dev).set_compute_dtype(dtype)
            teacher = teacher.train() if w.get("co_update") else teacher.eval()
        self.stepper = PretrainStepper(student, teacher, use_graphs=bool(args.graphs), tasks=("mlm", "sap"),
                                       co_update=bool(w.get("co_update")), side_stream=bool(args.side_stream),
                                       branch_streams=bool(args.branch_streams), overlap=bool(args.overlap),
                                       pipeline_teacher=bool(args.pipeline), teacher_sm_budget=args.teacher_sms,
                                       pdl=None if args.pdl < 0 else ((args.pdl & 1) != 0, (args.pdl & 2) != 0))
        ops.set_seed(dev, 1234 + rank)
        n = self.pool_n = args.pool
        store_cpu = None
        self.store = None
        if args.feature_store > 0:
            # GPU batch featuriser: the panorama features are resident in HBM (bf16 in bf16 mode), a batch names
            # panorama rows and view orders, and the model gathers its inputs on the device
            from magic_b200 import synth
            from magic_b200.featurizer import FeatureStore
            store_cpu = synth.make_store(args.feature_store, seed=77, dtype=dtype)
            self.store = FeatureStore(store_cpu, device=dev, dtype=dtype).attach(student, teacher)
        pools = {t: make_pool(t, n, w, 1234 + rank * 1000 + (0 if t == "mlm" else 500), store_cpu)
                 for t in ("mlm", "sap")}
        # flat batches: every tensor of a batch is a view into one buffer, so staging a batch is ONE copy
        self.dev_pools = {t: [flatten_batch(b, device=dev) for b in bs] for t, bs in pools.items()}
        self.pin_pools = {t: [flatten_batch(b, pin=True) for b in bs] for t, bs in pools.items()}
        self.in_bytes = sum(nbytes(b) for bs in pools.values() for b in bs) / (2 * n)
        self.host_loss = [torch.zeros(1).pin_memory() for _ in range(2)]
        self.loss_ev = [torch.cuda.Event() for _ in range(2)]

    def pick(self, i):
        return ("mlm" if i % 2 == 0 else "sap"), (i // 2) % self.pool_n

    def run_steps(self, n, first, from_host):
        stepper, out, nxt = self.stepper, None, None
        if from_host:
            task, j = self.pick(first)
            nxt = stepper.prefetch(task, self.pin_pools[task][j])
        for i in range(first, first + n):
            task, j = self.pick(i)
            if from_host:
                # end to end: every step's inputs start in pinned host memory; the copy of step i+1 is issued on the
                # copy stream before step i's loss is read back, as the reference's PrefetchLoader does
                b = nxt
                if i + 1 < first + n:
                    t2, j2 = self.pick(i + 1)
                    nxt = stepper.prefetch(t2, self.pin_pools[t2][j2])
            else:
                b = self.dev_pools[task][j]
            # a prefetching loader knows the next batch: announce it so a frozen teacher's forward of step i+1 can run
            # under step i's backward (PretrainStepper.step)
            ahead = None
            if i + 1 < first + n:
                t2, j2 = self.pick(i + 1)
                ahead = (t2, nxt if from_host else self.dev_pools[t2][j2])
            out = stepper.step(task, b, next=ahead)
            if from_host:
                # device -> host read of EVERY step's loss, one step late (async copy into pinned memory + event),
                # so the host queues step i+1 while step i runs instead of draining the GPU each step
                k = i % 2
                self.host_loss[k].copy_(out[0:1], non_blocking=True)
                self.loss_ev[k].record()
                if i > first:
                    self.loss_ev[1 - k].synchronize()
                    _ = self.host_loss[1 - k].item()
        if from_host:
            self.loss_ev[(first + n - 1) % 2].synchronize()
            _ = self.host_loss[(first + n - 1) % 2].item()
        return out

    def sync_all(self):
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(self, steps, from_host):
        """-> (ms for `steps` steps: CUDA events bracketed by barrier + synchronize, max over ranks), launches."""
        import torch.distributed as dist
        from magic_b200 import _lib
        c0 = _lib.COUNTERS["launches"]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.sync_all()
        t0 = time.perf_counter()
        e0.record()
        self.run_steps(steps, 100, from_host)
        e1.record()
        self.sync_all()
        ms = e0.elapsed_time(e1)
        if from_host:
            ms = max(ms, (time.perf_counter() - t0) * 1e3)
        t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), _lib.COUNTERS["launches"] - c0

    def profile_graph(self, nprof, serial=True):
        """Per-kernel durations of the REPLAYED graphs: the step graphs are re-captured with an external event-record
        node before and after every C-ABI call (`_lib.profile_start(graph=True)`), replayed `nprof` times per task,
        and every bracket is read with cudaEventElapsedTime after each replay.  Every rank runs the pass (the
        gradient exchange stays matched); rank 0 reports."""
        from magic_b200 import _lib, ops
        st = self.stepper
        saved = st.graphs, st.pipeline_teacher, st._t_inflight
        st.graphs, st._t_inflight = {}, None
        if serial:
            # one chain: no stream branches, no weight-gradient side stream, teacher inside the step's graph -- every
            # bracket then measures ITS kernel alone on the GPU and the brackets add up to the (serial) step
            st.pipeline_teacher = False
            ops.enable_branch_streams(False)
            ops.enable_side_stream(False)
        _lib.profile_start(graph=True)
        try:
            for i in range(2):  # capture (+ first replay) of the MLM and the SAP graph
                _lib.profile_tag(self.pick(i)[0])
                self.run_steps(1, i, False)
            torch.cuda.synchronize()
            prof = _lib._PROFILE
            acc = {name: [0.0] * len(recs) for name, recs in prof.items()}
            wall = 0.0
            for i in range(2 * nprof):
                task = self.pick(i)[0]
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                self.run_steps(1, 100 + i, False)
                e1.record()
                torch.cuda.synchronize()
                wall += e0.elapsed_time(e1)
                for name, recs in prof.items():
                    a = acc[name]
                    for k, r in enumerate(recs):
                        if r[3] == task:
                            a[k] += r[0].elapsed_time(r[1])
        finally:
            prof = _lib.profile_stop()
            st.graphs, st.pipeline_teacher, st._t_inflight = saved[0], saved[1], None
            ops.enable_branch_streams(bool(self.args.branch_streams))
            ops.enable_side_stream(bool(self.args.side_stream))
        # per record: mean ms per replay of ITS task's graph -> per step (MLM and SAP alternate: each runs nprof times)
        out = {name: [(a / nprof, r[2]) for a, r in zip(acc[name], recs)] for name, recs in prof.items()}
        return out, wall / (2 * nprof)


FAMILY_OF = {"magic_gemm_wgrad": "magic_gemm", "magic_attn_bwd_part": "magic_attn_bwd"}


def summarise_graph_profile(prof, ms_profiled_step, pk, workload):
    """prof: {name: [(ms summed over one MLM+SAP pair ... per record, args)]}; each record ran in ONE of the two task
    graphs, so a record's mean duration contributes half of it to the average step."""
    fam, shapes = {}, {}
    for name, recs in prof.items():
        key = FAMILY_OF.get(name, name)
        d = fam.setdefault(key, dict(ms_per_step=0.0, calls_per_step=0.0, flops=0.0, bytes=0.0))
        for ms, a in recs:
            f, b = family_cost(name, a)
            d["ms_per_step"] += ms / 2
            d["calls_per_step"] += 0.5
            d["flops"] += f / 2
            d["bytes"] += b / 2
            if key == "magic_gemm":
                # the same kernel serves the h = 768 encoder GEMMs (>= 1 GFLOP per launch: tensor-bound territory) and
                # the one-tile h = 128 student GEMMs (launch / latency bound): reported together AND split
                sub = fam.setdefault("magic_gemm_ge1gflop" if f >= 1e9 else "magic_gemm_lt1gflop",
                                     dict(ms_per_step=0.0, calls_per_step=0.0, flops=0.0, bytes=0.0, split=True))
                sub["ms_per_step"] += ms / 2
                sub["calls_per_step"] += 0.5
                sub["flops"] += f / 2
                sub["bytes"] += b / 2
                M, N, K = (a[11], a[12], a[13]) if name == "magic_gemm" else (a[10], a[11], a[9])
                sh = shapes.setdefault((name[6:], M, N, K), [0.0, 0, 0.0])
                sh[0] += ms / 2
                sh[1] += 1
                sh[2] += f / 2
    tot = sum(v["ms_per_step"] for v in fam.values() if not v.get("split")) or 1.0
    for v in fam.values():
        v["share"] = v["ms_per_step"] / tot

    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(workload, {})
    except Exception:
        traffic = {}

    def roof(names, bound, label):
        ms = sum(fam[n]["ms_per_step"] for n in names if n in fam)
        if ms <= 0:
            return None
        calls = sum(fam[n]["calls_per_step"] for n in names if n in fam)
        if bound == "tensor":
            work = sum(fam[n]["flops"] for n in names if n in fam)
            ach, peak, unit = work / (ms * 1e-3) / 1e12, pk["tf_sus"], "TFLOP/s"
            src = pk["src"] + " (sustained: timed inside a long step)"
        else:
            work = sum(fam[n]["bytes"] for n in names if n in fam)
            ach, peak, unit = work / (ms * 1e-3) / 1e9, pk["hbm"], "GB/s"
            src = pk["src"]
        r = dict(kernel=label, bound=bound, achieved=ach, peak=peak, unit=unit, frac=ach / peak, traffic=None,
                 peak_source=src, avg_launch_us=ms * 1e3 / max(calls, 1e-9), launches_per_step=calls,
                 ms_per_step=ms, share_of_step_kernel_time=ms / tot,
                 algorithmic_per_launch=work / max(calls, 1e-9))
        ent = traffic.get(label)
        if ent:
            r["traffic"] = ent["dram_bytes_per_launch"]
            r["traffic_source"] = ent["source"]
        return r

    rooflines = [r for r in (roof(["magic_gemm"], "tensor", "magic_gemm"),
                             roof(["magic_gemm_ge1gflop"], "tensor", "magic_gemm (launches >= 1 GFLOP: h=768 encoder GEMMs)"),
                             roof(["magic_gemm_lt1gflop"], "tensor", "magic_gemm (launches < 1 GFLOP: one-tile h=128 GEMMs)"),
                             roof(["magic_attn_fwd", "magic_attn_bwd"], "hbm", "magic_attn"),
                             roof(["magic_makd_mse_fwd", "magic_makd_mse_bwd", "magic_makd_kl_fwd",
                                   "magic_makd_kl_bwd"], "hbm", "magic_makd")) if r]
    top = max([r for r in rooflines if "(" not in r["kernel"]], key=lambda r: r["ms_per_step"]) if rooflines else None
    fams = {k: dict(ms_per_step=round(x["ms_per_step"], 4), calls=round(x["calls_per_step"], 1),
                    share=round(x["share"], 3),
                    tflops=round(x["flops"] / (x["ms_per_step"] * 1e-3) / 1e12, 2) if x["flops"] else None,
                    gbs=round(x["bytes"] / (x["ms_per_step"] * 1e-3) / 1e9, 1) if x["bytes"] else None)
            for k, x in sorted(((k, x) for k, x in fam.items() if not x.get("split")),
                               key=lambda kv: -kv[1]["ms_per_step"])[:10]}
    gshapes = [dict(op=k[0], M=k[1], N=k[2], K=k[3], launches_per_step=v[1] / 2, ms_per_step=round(v[0], 4),
                    tflops=round(v[2] / (v[0] * 1e-3) / 1e12, 1) if v[0] > 0 else None)
               for k, v in sorted(shapes.items(), key=lambda kv: -kv[1][0])[:12]]
    note = dict(mode="CUDA events inside a replayed graph of the same step captured on ONE stream (no branches): each "
                     "bracket times its kernel alone on the GPU", sum_kernel_ms=round(tot, 4),
                profiled_step_ms=round(ms_profiled_step, 4),
                note="the timed region replays the multi-branch graphs (value / ms_per_step); the sum of these "
                     "single-stream kernel times is the serial length of the same step")
    return top, rooflines, fams, gshapes, note


def run_ours(args):
    import torch.distributed as dist
    from magic_b200 import _lib
    from magic_b200.parallel import init_distributed
    rank, world, local = init_distributed()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    global WORKLOAD_NAME
    WORKLOAD_NAME = args.workload
    pk = peaks()
    R = Runner(args.workload, args, rank, world, dev)
    w = R.w
    # warm-up (also builds the CUDA graphs, one per task/shape)
    R.run_steps(max(args.warmup, 3), 0, False)
    R.sync_all()
    clk = ClockSampler(local)
    if rank == 0:
        clk.start()
    ms, launches = R.timed(args.steps, False)
    if args.timed_only:
        clk.stop() if rank == 0 else None
        sys.stderr.write(f"timed-only: {ms / args.steps:.3f} ms/step, {launches} launches\n")
        if world > 1:
            dist.destroy_process_group()
        return
    # end-to-end: host (pinned) buffers -> H2D -> step -> D2H loss, through the public stepper API
    if args.no_e2e:
        ms_e2e = ms
    else:
        R.run_steps(2, 0, True)
        ms_e2e, _ = R.timed(args.steps, True)
    clocks = clk.stop() if rank == 0 else None  # sampled (20 ms period) across BOTH timed regions
    B = w["B"]
    value = world * B * args.steps / (ms * 1e-3)
    e2e_v = world * B * args.steps / (ms_e2e * 1e-3)

    top = rooflines = fams = gshapes = note = None
    if args.graphs and not args.no_profile:
        try:
            prof, ms_prof = R.profile_graph(args.profile_steps, serial=bool(args.profile_serial))
            if rank == 0:
                top, rooflines, fams, gshapes, note = summarise_graph_profile(prof, ms_prof, pk, args.workload)
        except Exception as e:  # keep the headline numbers if the stopwatch graph cannot be built
            import traceback
            traceback.print_exc()
            note = dict(mode="unavailable", error=repr(e)[:300])
    line = None
    if rank == 0:
        cfg = workload_config(args.workload, w, world, args.dropout, R.stepper.use_graphs, R.pool_n, R.in_bytes)
        cfg["model_tflops"] = value * train_gflop_per_sample(w) / 1e3
        cfg["features"] = ("device-resident store of %d panoramas x 36 x 768 %s (%.0f MB); a batch carries panorama rows + "
                           "view orders and the model gathers its inputs in HBM" % (
                               R.store.N, args.dtype, R.store.nbytes / 1e6)) if R.store is not None else \
            "fp32 36 x 768 view features travel host->device every step (reference loader layout)"
        cfg["gradient_exchange"] = R.stepper.exchange_description() if world > 1 else None
        line = dict(
            metric=f"pretrain samples/s ({kind_of(w)})", value=value, unit="samples/s", n_gpus=world, steps=args.steps,
            warmup=max(args.warmup, 3), ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak",
            vs_baseline=None, dtype=args.dtype, data="synthetic", config=cfg,
            roofline=top, rooflines=rooflines, kernel_families=fams, gemm_shapes=gshapes, profile=note,
            e2e=dict(value=e2e_v, unit="samples/s", h2d_bytes_per_step=int(R.in_bytes), d2h_bytes_per_step=4,
                     ms_per_step=ms_e2e / args.steps),
            gpu_launches=int(launches), clocks=clocks)
    del R
    torch.cuda.empty_cache()

    # the other BASELINE.json configs that fit this GPU count, short runs (device-resident inputs, graph replay)
    subs = {}
    names = [] if args.sub_workloads == "none" else (
        [n for n in SUB_WORKLOADS if n != args.workload] if args.sub_workloads == "auto"
        else [n for n in args.sub_workloads.split(",") if n])
    for name in names:
        try:
            Rs = Runner(name, args, rank, world, dev)
            Rs.run_steps(4, 0, False)
            sms, sl = Rs.timed(args.sub_steps, False)
            ws = Rs.w
            v = world * ws["B"] * args.sub_steps / (sms * 1e-3)
            subs[name] = dict(metric=f"pretrain samples/s ({kind_of(ws)})", value=v, unit="samples/s",
                              ms_per_step=sms / args.sub_steps, steps=args.sub_steps, n_gpus=world,
                              gpu_launches=int(sl), model_tflops=v * train_gflop_per_sample(ws) / 1e3,
                              config=workload_config(name, ws, world, args.dropout, Rs.stepper.use_graphs))
            del Rs
        except Exception as e:
            subs[name] = dict(error=repr(e)[:300])
        torch.cuda.empty_cache()
    if rank == 0:
        line["workloads"] = subs or None
        line["cpu_baseline"] = line["gpu_baseline"] = None
        if world == 1 and not args.no_gpu_baseline:
            try:
                v, gms = torch_gpu_rate(w, args.dropout)
                line["gpu_baseline"] = dict(value=v, unit="samples/s", ms_per_step=gms,
                                            kind="oracle port, stock torch ops on cuda, bf16 autocast, eager, "
                                                 "fused torch AdamW (bench.py --impl torch_gpu)")
            except Exception as e:
                line["gpu_baseline"] = dict(error=repr(e)[:300])
            torch.cuda.empty_cache()
        if world == 1 and not args.no_extras:
            for key, fn in (("nav_inference", nav_latency), ("featurizer", featurizer_rate)):
                try:
                    line[key] = fn()
                except Exception as e:
                    line[key] = dict(error=repr(e)[:300])
                torch.cuda.empty_cache()
        if world == 1 and not args.no_cpu:
            v, n, tstep = cpu_step_rate(w, args.cpu_seconds, args.dropout)
            line["cpu_baseline"] = dict(
                value=v, unit="samples/s", cores=os.cpu_count(), kind="port",
                sample=f"fp32 PyTorch oracle port of the same step at the same batch size {B} "
                       f"(teacher forward + student forward/backward + MAKD + clip + AdamW), MLM/SAP 1:1, {n} steps "
                       f"on the host cores, mean {tstep * 1e3:.0f} ms/step after the first")
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
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_gpu"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--sub-workloads", default="auto",
                    help="'auto' = the other BASELINE.json configs, 'none', or a comma-separated list")
    ap.add_argument("--sub-steps", type=int, default=10)
    ap.add_argument("--profile-steps", type=int, default=3)
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--profile-serial", type=int, default=1,
                    help="1: the stopwatch graph is captured on one stream (kernels timed alone); 0: same branches as the timed graphs")
    ap.add_argument("--no-e2e", action="store_true", help="debug aid: skip the end-to-end region")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--compile", type=int, default=0, help="--impl torch_gpu: wrap the step in torch.compile")
    ap.add_argument("--overlap", type=int, default=1, help="N > 1: exchange gradient buckets during backward")
    ap.add_argument("--feature-store", type=int, default=2048,
                    help="panoramas in the device-resident feature store (0: batches carry fp32 features, as the "
                         "reference loader ships them)")
    ap.add_argument("--pdl", type=int, default=-1,
                    help="programmatic dependent launch: -1 stepper default, bit 0 = student graph, bit 1 = teacher graph")
    ap.add_argument("--teacher-sms", type=int, default=0,
                    help="SMs the pipelined teacher graph's persistent GEMMs may occupy (0 = all)")
    ap.add_argument("--pipeline", type=int, default=1,
                    help="frozen teacher: run the next batch's teacher forward under this batch's student backward")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--graphs", type=int, default=1)
    ap.add_argument("--pool", type=int, default=8)
    ap.add_argument("--side-stream", type=int, default=1)
    ap.add_argument("--branch-streams", type=int, default=1)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the nav-inference / featuriser side measurements")
    ap.add_argument("--timed-only", action="store_true",
                    help="profiling aid: run warm-up + the timed region only (no e2e / roofline / CPU passes, no JSON)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "torch_gpu":
        run_torch_gpu(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
