"""Bisect of a stream-capture invalidation: python scripts/capture_repro.py <variant>"""
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import magic_b200  # noqa: E402
from magic_b200 import _lib, ops  # noqa: E402
from magic_b200.graph_index import flatten_batch  # noqa: E402
from magic_b200.train_step import PretrainStepper  # noqa: E402

variant = sys.argv[1]
dev = torch.device("cuda", 0)
w = dict(bench.WORKLOADS["magic_s_distill_t768_b64"])
w["B"] = 8
w["teacher"] = dict(hidden=256, n_l=6, n_x=3, n_p=2)
cfg_s, cfg_t = bench.make_cfgs(w, 0.1)
torch.manual_seed(1)
student = magic_b200.GlocalTextPathCMTPreTraining(cfg_s).to(dev).train().set_compute_dtype(torch.bfloat16)
teacher = magic_b200.GlocalTextPathCMTPreTraining(cfg_t).to(dev).eval().set_compute_dtype(torch.bfloat16)
pools = {t: [flatten_batch(b, device=dev) for b in bench.make_pool(t, 2, w, 7 + (0 if t == "mlm" else 50))]
         for t in ("mlm", "sap")}


def steps(st, n, first=0):
    for i in range(first, first + n):
        task = "mlm" if i % 2 == 0 else "sap"
        st.step(task, pools[task][(i // 2) % 2])
    torch.cuda.synchronize()


def serial(st):
    st.graphs, st._t_inflight = {}, None
    st.pipeline_teacher = False
    ops.enable_branch_streams(False)
    ops.enable_side_stream(False)


try:
    st = PretrainStepper(student, teacher, use_graphs=True, pipeline_teacher=variant not in ("v1", "v5"))
    ops.set_seed(dev, 1)
    if variant == "v1":      # serial from the start
        serial(st)
        steps(st, 4)
    elif variant == "v2":    # branchy graphs first, then a serial re-capture
        steps(st, 4)
        serial(st)
        steps(st, 4)
    elif variant == "v3":    # as v2, with the stopwatch events
        steps(st, 4)
        serial(st)
        _lib.profile_start(graph=True)
        steps(st, 4)
        _lib.profile_stop()
    elif variant == "v4":    # branchy re-capture with events (the mode that worked before)
        steps(st, 4)
        st.graphs, st._t_inflight = {}, None
        _lib.profile_start(graph=True)
        steps(st, 4)
        _lib.profile_stop()
    elif variant == "v5":    # non-pipelined branchy, then serial
        steps(st, 4)
        serial(st)
        steps(st, 4)
    elif variant == "v6":    # as v2 but keep the side stream on (only branches off)
        steps(st, 4)
        st.graphs, st._t_inflight = {}, None
        st.pipeline_teacher = False
        ops.enable_branch_streams(False)
        steps(st, 4)
    elif variant == "v7":    # as v2 but keep the branches (only the side stream off)
        steps(st, 4)
        st.graphs, st._t_inflight = {}, None
        st.pipeline_teacher = False
        ops.enable_side_stream(False)
        steps(st, 4)
    print(variant, "OK")
except Exception:
    traceback.print_exc()
    print(variant, "FAILED")
