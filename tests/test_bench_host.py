"""CPU tests of bench.py's host-side bookkeeping (no GPU): per-family roofline summary with outlier clipping and the
traffic lookup, the one-line stdout contract helpers, and ops.GradLink's hand-over semantics."""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _Ev:
    def __init__(self, t):
        self.t = t

    def elapsed_time(self, other):
        return other.t - self.t


def _rec(ms, args):
    return (_Ev(0.0), _Ev(ms), args)


def test_summarise_profile_clips_one_off_stalls_and_reports_traffic():
    bench = importlib.import_module("bench")
    bench.OUTLIERS.clear()
    bench.WORKLOAD_NAME = "magic_s_pretrain_b64"
    # magic_ln_fwd args: (..., M at [6], h at [7], ..., dtype at [9]); 40 normal calls and one 39 ms stall
    ln_args = (0, 0, 0, 0, 0, 0, 5120, 128, 1e-12, 1)
    prof = {"magic_ln_fwd": [_rec(0.005, ln_args) for _ in range(40)] + [_rec(39.0, ln_args)],
            "magic_delay": [_rec(1.0, ())]}
    pk = dict(hbm=6546.2, tf=1661.3, tf_sus=1403.5, src="measured")
    roof, fams = bench.summarise_profile(prof, 1, pk)
    assert bench.OUTLIERS == {"magic_ln_fwd": 1}
    assert roof["kernel"] == "magic_ln_fwd" and roof["bound"] == "hbm"
    assert abs(fams["magic_ln_fwd"]["ms_per_step"] - 41 * 0.005) < 1e-6  # the stall counts as the median
    by = 41 * 5120 * 128 * 2 * 3
    assert abs(roof["achieved"] - by / (41 * 0.005e-3) / 1e9) / roof["achieved"] < 1e-6
    assert roof["peak"] == 6546.2 and 0 < roof["frac"] < 1
    # GEMM family: traffic comes from the committed ncu capture
    g_args = (0, 1, 128, 1, 0, 1, 1, 128, 0, 1, 128, 5120, 128, 128)
    prof = {"magic_gemm": [_rec(0.004, g_args) for _ in range(10)]}
    roof, _ = bench.summarise_profile(prof, 1, pk)
    tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    assert roof["bound"] == "tensor" and roof["traffic"] == tr["magic_s_pretrain_b64"]["magic_gemm"]["dram_bytes_per_launch"]
    assert abs(roof["achieved"] - 2.0 * 5120 * 128 * 128 / 0.004e-3 / 1e12) < 1e-6


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
