"""Probe: how far are bf16 gradients from the fp32 oracle -- for our bf16 path and for stock torch autocast."""
import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_model_gpu import build_pair, get_batch, oracle_batch, rel

task = sys.argv[1] if len(sys.argv) > 1 else "sap"
oracle, prod = build_pair(128, ht=256, seed=5)
prod.set_compute_dtype(torch.bfloat16).train()
oracle.train()
b = get_batch(task, seed=5)
oracle(oracle_batch(b), task, True)["loss"].mean().backward()
ref = {n: p.grad.clone() for n, p in oracle.named_parameters() if p.grad is not None}
oracle.zero_grad()
with torch.autocast("cuda", dtype=torch.bfloat16):
    l = oracle(oracle_batch(b), task, True)["loss"].float().mean()
l.backward()
ac = {n: p.grad.clone() for n, p in oracle.named_parameters() if p.grad is not None}
prod(b, task, True)["loss"].mean().backward()
mine = {n: p.grad for n, p in prod.named_parameters() if p.grad is not None}
gmax = max(g.norm().item() for g in ref.values())
rows = []
for n, g in ref.items():
    if g.norm() < 1e-4 * gmax:
        continue
    rows.append((rel(mine[n], g), rel(ac[n], g), n))
rows.sort(reverse=True)
for r in rows[:25]:
    print("mine %.3f  autocast %.3f  %s" % r)
print("median mine %.3f autocast %.3f" % (sorted(r[0] for r in rows)[len(rows)//2], sorted(r[1] for r in rows)[len(rows)//2]))
