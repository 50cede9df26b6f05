"""Generates tests/golden/adamw_tasks_ref.pt with the REFERENCE's optimizer (optim/adamw.py, misc.py, sched.py) for the
case a multi-task loop produces: two heads that belong to different tasks, so in every step the parameters of the
OTHER task's head have `p.grad is None` -- the reference skips them (adamw.py:66-67: no moment decay, no weight decay)
and counts their steps separately (:86 state['step']).
Run:  python tests/golden/gen_adamw_tasks_golden.py      (needs /root/reference; the output is committed)"""
import os
import sys
from types import SimpleNamespace

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/pretrain_src")


class TwoHeads(nn.Module):
    def __init__(self):
        super().__init__()
        self.trunk = nn.Linear(24, 40)
        self.LayerNorm = nn.LayerNorm(40)
        self.head_a = nn.Linear(40, 5)
        self.head_b = nn.Linear(40, 7)
        self.never = nn.Linear(8, 8)   # a head no configured task touches (e.g. an unused KD projection)


TASKS = ["a", "b", "a", "a", "b", "b", "a", "b"]


def inactive(task, name):
    return name.startswith("never.") or name.startswith("head_b." if task == "a" else "head_a.")


def main():
    from optim.misc import build_optimizer
    from optim.sched import get_lr_sched
    torch.manual_seed(20261018)
    model = TwoHeads()
    opts = SimpleNamespace(optim="adamw", learning_rate=5e-5, betas=[0.9, 0.98], weight_decay=0.01, warmup_steps=3,
                           num_train_steps=12, grad_norm=5.0)
    opt = build_optimizer(model, opts)
    g = torch.Generator().manual_seed(11)
    init = {n: p.detach().clone() for n, p in model.named_parameters()}
    steps = []
    for step, task in enumerate(TASKS, 1):
        lr = get_lr_sched(step, opts)
        for pg in opt.param_groups:
            pg["lr"] = lr
        grads = {}
        scale = 30.0 if step in (3, 6) else 0.05
        for n, p in model.named_parameters():
            if inactive(task, n):
                p.grad = None
                continue
            grads[n] = torch.randn(p.shape, generator=g) * scale
            p.grad = grads[n].clone()
        norm = torch.nn.utils.clip_grad_norm_([p for p in model.parameters() if p.grad is not None], opts.grad_norm)
        opt.step()
        steps.append(dict(step=step, task=task, lr=lr, grads=grads, grad_norm=float(norm),
                          params={n: p.detach().clone() for n, p in model.named_parameters()}))
    torch.save(dict(init=init, steps=steps, opts=vars(opts), tasks=TASKS), os.path.join(HERE, "adamw_tasks_ref.pt"))
    print("wrote adamw_tasks_ref.pt:", len(steps), "steps")


if __name__ == "__main__":
    main()
