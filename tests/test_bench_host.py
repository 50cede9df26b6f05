"""CPU tests of bench.py's host-side bookkeeping (no GPU): per-family roofline summary with outlier clipping and the
traffic lookup, the one-line stdout contract helpers, and ops.GradLink's hand-over semantics."""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_graph_profile_summary_three_rooflines_and_shapes():
    """bench.summarise_graph_profile: records are (mean ms per replay of the record's task graph, C-ABI args); MLM and
    SAP graphs alternate, so each record weighs one half in the average step."""
    bench = importlib.import_module("bench")
    pk = dict(hbm=6546.2, tf=1661.3, tf_sus=1403.5, src="measured")
    # magic_gemm args: M, N, K at [11], [12], [13]; dtypes at [1], [5], [9]
    g_args = (0, 1, 768, 1, 0, 1, 1, 768, 0, 1, 768, 5120, 768, 768)
    # magic_gemm_wgrad args: (dy, dy_dt, dy_ld, x, x_dt, x_ld, dw, dw_ld, dbias, M, N, K)
    w_args = (0, 1, 768, 0, 1, 768, 0, 768, 0, 5120, 768, 768)
    # magic_attn_fwd: B, H, Lq, Lk at [11..14], dtype at [20]
    a_args = (0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 64, 12, 80, 80, 0, 0, 0, 0, 0, 1)
    prof = {"magic_gemm": [(0.010, g_args)] * 20, "magic_gemm_wgrad": [(0.012, w_args)] * 4,
            "magic_attn_fwd": [(0.008, a_args)] * 10}
    top, roofs, fams, shapes, note = bench.summarise_graph_profile(prof, 0.2, pk, "no_such_workload")
    assert top["kernel"] == "magic_gemm" and top["bound"] == "tensor" and top["traffic"] is None
    gemm_ms = (20 * 0.010 + 4 * 0.012) / 2
    assert abs(fams["magic_gemm"]["ms_per_step"] - gemm_ms) < 1e-4
    fl = 2.0 * 5120 * 768 * 768 * 24 / 2
    assert abs(top["achieved"] - fl / (gemm_ms * 1e-3) / 1e12) / top["achieved"] < 1e-6
    assert abs(top["frac"] - top["achieved"] / 1403.5) < 1e-9
    assert abs(top["avg_launch_us"] - gemm_ms * 1e3 / 12) < 1e-6
    kinds = {r["kernel"]: r for r in roofs}
    big = "magic_gemm (launches >= 1 GFLOP: h=768 encoder GEMMs)"   # the family is also reported split by launch size
    assert set(kinds) == {"magic_gemm", big, "magic_attn"} and kinds["magic_attn"]["bound"] == "hbm"
    assert abs(kinds[big]["achieved"] - top["achieved"]) / top["achieved"] < 1e-9   # (every launch here is 6 GFLOP)
    assert kinds["magic_attn"]["peak"] == 6546.2
    assert {(d["op"], d["M"], d["N"], d["K"]) for d in shapes} == {("gemm", 5120, 768, 768), ("gemm_wgrad", 768, 768, 5120)}
    assert abs(note["sum_kernel_ms"] - (gemm_ms + 10 * 0.008 / 2)) < 1e-4 and note["profiled_step_ms"] == 0.2
    # the traffic of a roofline comes from the committed ncu capture of THAT workload, when there is one
    tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    for wl, ents in tr.items():
        for label, ent in ents.items():
            fam_prof = {"magic_gemm": [(0.010, g_args)]} if label == "magic_gemm" else None
            if fam_prof:
                t2 = bench.summarise_graph_profile(fam_prof, 0.01, pk, wl)[0]
                assert t2["traffic"] == ent["dram_bytes_per_launch"]


def test_train_gflop_accounting():
    bench = importlib.import_module("bench")
    w = bench.WORKLOADS["magic_s_pretrain_b64"]
    assert abs(bench.train_gflop_per_sample(w) - 3 * 0.669) < 1e-9
    w = bench.WORKLOADS["magic_s_distill_t768_b64"]
    assert abs(bench.train_gflop_per_sample(w) - (3 * 0.669 + 22.08)) < 1e-9      # frozen teacher: forward only
    w = bench.WORKLOADS["magic_l_icod_b32"]
    assert abs(bench.train_gflop_per_sample(w) - (3 * 17.29 + 3 * 22.08)) < 1e-9  # ICoD: both models train


def test_grad_link_hands_the_gradient_over_once():
    import magic_b200  # noqa: F401  (package import must not need the GPU)
    from magic_b200 import ops
    link = ops.GradLink()
    like = torch.zeros(6, 4)
    assert link.take(like) is None
    link.dres = torch.arange(24.0).reshape(2, 3, 4)  # LayerNorm parks a [B, L, h] view
    got = link.take(like)
    assert got.shape == like.shape and got.is_contiguous() and torch.equal(got.reshape(-1), torch.arange(24.0))
    assert link.dres is None and link.take(like) is None  # consumed exactly once
