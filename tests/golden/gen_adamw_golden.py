"""Generates tests/golden/adamw_ref.pt by running the REFERENCE's own optimizer files in the build container:
  /root/reference/pretrain_src/optim/adamw.py   (AdamW, eps 1e-6, bias correction, decoupled decay after the update)
  /root/reference/pretrain_src/optim/misc.py    (build_optimizer: no_decay grouping on bias / LayerNorm.*)
  /root/reference/pretrain_src/optim/sched.py   (get_lr_sched: warm-up-linear)
on a small named parameter set for 6 steps of the (missing) loop's optimizer phase, in the order the config implies
(r2r_magic_pretrain.json:22 grad_norm 5.0): lr = get_lr_sched(step); clip_grad_norm_(5.0); optimizer.step().
Run:  python tests/golden/gen_adamw_golden.py      (needs /root/reference; the output is committed)"""
import os
import sys
from types import SimpleNamespace

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/pretrain_src")


class Tiny(nn.Module):
    """Parameter names exercise the grouping rule: *.bias and LayerNorm.* get no weight decay."""

    def __init__(self):
        super().__init__()
        self.dense = nn.Linear(24, 40)
        self.LayerNorm = nn.LayerNorm(40)
        self.emb = nn.Embedding(17, 8)
        self.out = nn.Linear(40, 3, bias=False)


def main():
    from optim.misc import build_optimizer
    from optim.sched import get_lr_sched
    torch.manual_seed(20261017)
    model = Tiny()
    opts = SimpleNamespace(optim="adamw", learning_rate=5e-5, betas=[0.9, 0.98], weight_decay=0.01, warmup_steps=4,
                           num_train_steps=10, grad_norm=5.0)
    opt = build_optimizer(model, opts)
    g = torch.Generator().manual_seed(7)
    init = {n: p.detach().clone() for n, p in model.named_parameters()}
    steps = []
    for step in range(1, 7):
        lr = get_lr_sched(step, opts)
        for pg in opt.param_groups:
            pg["lr"] = lr
        grads = {}
        scale = 30.0 if step in (2, 5) else 0.05  # two steps clip, the others do not
        for n, p in model.named_parameters():
            grads[n] = torch.randn(p.shape, generator=g) * scale
            p.grad = grads[n].clone()
        norm = torch.nn.utils.clip_grad_norm_(model.parameters(), opts.grad_norm)
        opt.step()
        steps.append(dict(step=step, lr=lr, grads=grads, grad_norm=float(norm),
                          params={n: p.detach().clone() for n, p in model.named_parameters()}))
    sched = [(s, get_lr_sched(s, opts)) for s in range(0, 13)]
    torch.save(dict(init=init, steps=steps, sched=sched, opts=vars(opts),
                    no_decay=[n for n, _ in model.named_parameters()
                              if any(nd in n for nd in ("bias", "LayerNorm.bias", "LayerNorm.weight"))]),
               os.path.join(HERE, "adamw_ref.pt"))
    print("wrote adamw_ref.pt:", len(steps), "steps; clipped at", [s["step"] for s in steps if s["grad_norm"] > 5.0])


if __name__ == "__main__":
    main()
