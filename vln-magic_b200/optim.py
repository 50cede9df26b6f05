"""Fused optimizer for the flat arena: global-norm clip + AdamW + bf16 shadow refresh in two launches per
group, with the reference's update order and hyper-parameters (pretrain_src/optim/adamw.py:53-112 -- eps 1e-6,
bias correction, decoupled weight decay applied AFTER the Adam update with lr; grouping optim/misc.py:13-22;
schedule optim/sched.py:17-30; clip `grad_norm` r2r_magic_pretrain.json:22).  Hyper-parameters that change
per step live in a device buffer so a captured CUDA graph can replay the step unchanged."""
import math

import torch

from ._lib import call, ptr, stream
from .arena import ParamArena
from .ops import PinnedRing

SUMSQ_SCRATCH = 2048  # MAGIC_SUMSQ_SCRATCH (include/magic_b200.h)


def warmup_linear(step, warmup_step, tot_step):
    if step < warmup_step:
        return step / warmup_step
    return max(0, (tot_step - step) / (tot_step - warmup_step))


def get_lr_sched(global_step, opts):
    lr = opts.learning_rate * warmup_linear(global_step, opts.warmup_steps, opts.num_train_steps)
    return lr if lr > 0 else 1e-8


class FusedAdamW:
    def __init__(self, arena: ParamArena, lr=5e-5, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.01, max_grad_norm=5.0):
        self.arena = arena
        self.lr, self.betas, self.eps, self.weight_decay, self.max_grad_norm = lr, tuple(betas), eps, weight_decay, \
            max_grad_norm
        dev = arena.device
        self.m = torch.zeros(arena.total, dtype=torch.float32, device=dev)
        self.v = torch.zeros(arena.total, dtype=torch.float32, device=dev)
        self.hyper = torch.zeros(8, dtype=torch.float32, device=dev)
        self.hyper_ring = PinnedRing(8)
        self.sumsq = torch.zeros(1 + SUMSQ_SCRATCH, dtype=torch.float32, device=dev)  # [0] = sum g^2, rest = scratch
        self.step_count = 0
        self.param_groups = [dict(lr=lr)]  # so `for g in optimizer.param_groups: g['lr'] = ...` keeps working

    def set_hyper(self, lr=None):
        """Host -> device copy of the per-step scalars (call OUTSIDE a captured graph)."""
        self.step_count += 1
        lr = self.param_groups[0]["lr"] if lr is None else lr
        b1, b2 = self.betas
        bc1, bc2 = 1.0 - b1 ** self.step_count, 1.0 - b2 ** self.step_count
        self.hyper_ring.upload([lr, lr * math.sqrt(bc2) / bc1, b1, b2, self.eps,
                                (self.max_grad_norm if self.max_grad_norm else 0.0), 0.0, 0.0], self.hyper)

    def apply(self):
        """Device work of one step (graph-capturable): grad-norm, AdamW on both groups, bf16 shadow."""
        a = self.arena
        st = stream()
        call("magic_sumsq", ptr(a.flat_g), a.total, ptr(self.sumsq), 1, st)
        shadow = a.flat_lowp
        nd = a.n_decay
        if nd > 0:
            call("magic_adamw", ptr(a.flat_p), ptr(a.flat_g), ptr(self.m), ptr(self.v), ptr(shadow), nd,
                 ptr(self.hyper), self.weight_decay, ptr(self.sumsq), st)
        rest = a.total - nd
        if rest > 0:
            o4, o2 = nd * 4, nd * 2
            call("magic_adamw", a.flat_p.data_ptr() + o4, a.flat_g.data_ptr() + o4, self.m.data_ptr() + o4,
                 self.v.data_ptr() + o4, (shadow.data_ptr() + o2) if shadow is not None else None, rest,
                 ptr(self.hyper), 0.0, ptr(self.sumsq), st)

    def step(self, lr=None):
        self.set_hyper(lr)
        self.apply()

    def zero_grad(self):
        self.arena.zero_grad()

    def state_dict(self):
        """{step, exp_avg, exp_avg_sq} on the CPU, flat in arena order, plus the layout that makes them meaningful
        (ModelSaver dumps this next to the weights, utils/save.py:41-45)."""
        return dict(step=self.step_count, exp_avg=self.m.detach().cpu().clone(), exp_avg_sq=self.v.detach().cpu().clone(),
                    layout=[(n, o, k) for n, _, o, k in self.arena.entries], lr=self.param_groups[0]["lr"])

    def load_state_dict(self, sd):
        if [(n, o, k) for n, _, o, k in self.arena.entries] != [tuple(x) for x in sd["layout"]]:
            raise ValueError("optimizer state was saved for a different parameter layout")
        self.step_count = int(sd["step"])
        self.m.copy_(sd["exp_avg"])
        self.v.copy_(sd["exp_avg_sq"])
        self.param_groups[0]["lr"] = sd.get("lr", self.lr)

    def grad_norm(self):
        return float(self.sumsq[0].sqrt())


def build_optimizer(model, opts, lowp=None):
    """Mirror of pretrain_src/optim/misc.py:12-37 for optim == 'adamw'."""
    if getattr(opts, "optim", "adamw") != "adamw":
        raise ValueError("invalid optimizer (the fused path implements the configured 'adamw')")
    lowp = (getattr(model, "compute_dtype", torch.float32) == torch.bfloat16) if lowp is None else lowp
    arena = ParamArena(model, lowp=lowp)
    return FusedAdamW(arena, lr=opts.learning_rate, betas=tuple(opts.betas), weight_decay=opts.weight_decay,
                      max_grad_norm=getattr(opts, "grad_norm", 5.0))
