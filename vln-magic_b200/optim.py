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


def task_tables(entries, n_decay, total, tasks, inactive, max_slots=16, max_segs=512):
    """Segment tables of the task-aware AdamW (magic_adamw_seg) from the arena layout: `entries` = [(name, offset,
    numel)] in arena order.  -> (slots, {task: [(bounds, codes) for the decay group, (bounds, codes) for the no-decay
    group]}): `slots[i]` = the set of tasks that step hyper slot i (slot 0 = every task); per task and group the sorted
    segment bounds (relative to the group's start, first 0, last the group's size) and one code per segment: the hyper
    slot of the parameters in it, or -1 when the task leaves them untouched.  Host-only, exact integer logic."""
    all_t = frozenset(tasks)
    owners = [frozenset(t for t in tasks if not inactive(t, n)) for n, _, _ in entries]
    slots = [all_t] + sorted({o for o in owners if o and o != all_t}, key=sorted)
    if len(slots) > max_slots:
        raise ValueError("too many distinct task-activity sets")
    slot_of = {o: i for i, o in enumerate(slots)}
    tables = {}
    for t in tasks:
        groups = []
        for g_lo, g_hi in ((0, n_decay), (n_decay, total)):
            bounds, codes = [0], []
            for (n, o, k), own in zip(entries, owners):
                if not (g_lo <= o < g_hi):
                    continue
                code = slot_of[own] if t in own else -1
                if codes and codes[-1] == code:
                    continue           # same code as the running segment: it simply extends
                if codes:
                    bounds.append(o - g_lo)
                codes.append(code)
            bounds.append(g_hi - g_lo)
            if not codes:
                codes = [0]
            if len(codes) > max_segs:
                raise ValueError("too many activity segments in the arena")
            groups.append((bounds, codes))
        tables[t] = groups
    return slots, tables


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
        # task activity (configure_tasks): parameters a task never touches are skipped in that task's steps and
        # count their own steps, like the reference (adamw.py:66-67 `if p.grad is None: continue`, :86 state['step'])
        self.slots = None          # [frozenset(tasks)] per hyper slot; slot 0 = every task
        self.slot_steps = None
        self.task_tables = {}      # task -> ((bounds, codes, nseg) for the decay group, same for the no-decay group)

    MAX_SLOTS = 16

    def configure_tasks(self, tasks, inactive):
        """`tasks`: the tasks the loop alternates between; `inactive(task, param_name) -> bool`: True when that
        parameter receives NO gradient in a step of that task (model.inactive_in_task).  Afterwards pass the task to
        set_hyper / apply / step."""
        a = self.arena
        slots, tables = task_tables([(n, o, k) for n, _, o, k in a.entries], a.n_decay, a.total, list(tasks), inactive,
                                    self.MAX_SLOTS)
        self.slots, self.slot_steps = slots, [0] * len(slots)
        self.hyper = torch.zeros(8 * self.MAX_SLOTS, dtype=torch.float32, device=a.device)
        self.hyper_ring = PinnedRing(8 * self.MAX_SLOTS)
        dev = a.device
        self.task_tables = {t: tuple((torch.tensor(bounds, dtype=torch.int32, device=dev),
                                      torch.tensor(codes, dtype=torch.int32, device=dev), len(codes), codes == [0])
                                     for bounds, codes in groups) for t, groups in tables.items()}
        return self

    def set_hyper(self, lr=None, task=None):
        """Host -> device copy of the per-step scalars (call OUTSIDE a captured graph)."""
        self.step_count += 1
        lr = self.param_groups[0]["lr"] if lr is None else lr
        b1, b2 = self.betas

        def row(t):
            bc1, bc2 = 1.0 - b1 ** t, 1.0 - b2 ** t
            return [lr, lr * math.sqrt(bc2) / bc1, b1, b2, self.eps, (self.max_grad_norm if self.max_grad_norm else 0.0),
                    0.0, 0.0]

        if self.slots is None or task is None or task not in self.task_tables:
            if self.slots is not None:  # configured, but this step is not attributed to a task: everything steps
                self.slot_steps = [c + 1 for c in self.slot_steps]
                vals = sum((row(max(c, 1)) for c in self.slot_steps), [])
                vals[:8] = row(self.step_count)
                self.hyper_ring.upload(vals + [0.0] * (8 * self.MAX_SLOTS - len(vals)), self.hyper)
                return
            self.hyper_ring.upload(row(self.step_count), self.hyper)
            return
        for i, own in enumerate(self.slots):
            if task in own:
                self.slot_steps[i] += 1
        vals = sum((row(max(c, 1)) for c in self.slot_steps), [])
        self.hyper_ring.upload(vals + [0.0] * (8 * self.MAX_SLOTS - len(vals)), self.hyper)

    def apply(self, task=None):
        """Device work of one step (graph-capturable): grad-norm, AdamW on both groups, bf16 shadow.  With
        configure_tasks() and a task, only the parameters that task touches are updated."""
        a = self.arena
        st = stream()
        call("magic_sumsq", ptr(a.flat_g), a.total, ptr(self.sumsq), 1, st)
        shadow = a.flat_lowp
        nd = a.n_decay
        tabs = self.task_tables.get(task) if task is not None else None
        for gi, (lo, n, wd) in enumerate(((0, nd, self.weight_decay), (nd, a.total - nd, 0.0))):
            if n <= 0:
                continue
            sh = (shadow.data_ptr() + lo * 2) if shadow is not None else None
            args = (a.flat_p.data_ptr() + lo * 4, a.flat_g.data_ptr() + lo * 4, self.m.data_ptr() + lo * 4,
                    self.v.data_ptr() + lo * 4, sh, n, ptr(self.hyper), wd, ptr(self.sumsq))
            if tabs is None or tabs[gi][3]:  # no table, or one segment on slot 0: the plain kernel
                call("magic_adamw", *args, st)
            else:
                bounds, codes, nseg, _ = tabs[gi]
                call("magic_adamw_seg", *args, ptr(bounds), ptr(codes), nseg, st)

    def step(self, lr=None, task=None):
        self.set_hyper(lr, task)
        self.apply(task)

    def zero_grad(self):
        self.arena.zero_grad()

    def state_dict(self):
        """{step, exp_avg, exp_avg_sq} on the CPU, flat in arena order, plus the layout that makes them meaningful
        (ModelSaver dumps this next to the weights, utils/save.py:41-45)."""
        return dict(step=self.step_count, exp_avg=self.m.detach().cpu().clone(), exp_avg_sq=self.v.detach().cpu().clone(),
                    layout=[(n, o, k) for n, _, o, k in self.arena.entries], lr=self.param_groups[0]["lr"],
                    slot_steps=list(self.slot_steps) if self.slot_steps is not None else None)

    def load_state_dict(self, sd):
        if [(n, o, k) for n, _, o, k in self.arena.entries] != [tuple(x) for x in sd["layout"]]:
            raise ValueError("optimizer state was saved for a different parameter layout")
        self.step_count = int(sd["step"])
        self.m.copy_(sd["exp_avg"])
        self.v.copy_(sd["exp_avg_sq"])
        self.param_groups[0]["lr"] = sd.get("lr", self.lr)
        if self.slot_steps is not None and sd.get("slot_steps") is not None and len(sd["slot_steps"]) == len(self.slot_steps):
            self.slot_steps = [int(x) for x in sd["slot_steps"]]

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
