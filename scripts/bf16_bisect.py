"""Where does the bf16 path lose gradient accuracy?  Runs the SAME product model twice (fp32 mode = the parity
reference, 1e-4 from the oracle; bf16 mode = the benchmarked path) and compares the gradients of the activations
along the backward chain: task logits -> cross-modal layers (global / local) -> their inputs -> text output.
  python scripts/bf16_bisect.py [sap|mlm]"""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import magic_b200  # noqa: E402,F401
from magic_b200 import model as M  # noqa: E402
from test_model_gpu import build_pair, get_batch  # noqa: E402

task = sys.argv[1] if len(sys.argv) > 1 else "sap"
oracle, prod32 = build_pair(128, ht=256, seed=5)
prod16 = copy.deepcopy(prod32).set_compute_dtype(torch.bfloat16)
b = get_batch(task, seed=5)


def run(prod):
    prod.train()
    grads, acts = {}, {}

    def tap(name, t):
        if torch.is_tensor(t) and t.requires_grad:
            acts[name] = t.detach().float()
            t.register_hook(lambda g, n=name: grads.__setitem__(n, g.detach().float().clone()))
        return t

    orig_cross, orig_attn, orig_ffn = M._cross_encoder, M._attn_block, M._ffn_block

    def cross(enc, x, ctx, B, Lx, Lc, x_lens, c_lens, fc, dists=None, sprel=None):
        tag = "global" if enc is prod.bert.global_encoder.encoder else "local"
        tap(f"{tag}.in", x)
        tap(f"{tag}.ctx", ctx)
        attns = []
        for i, layer in enumerate(enc.crossattention):
            a, p_self = orig_attn(layer.attention, x, None, B, Lx, Lx, x_lens, fc, dists, sprel)
            tap(f"{tag}.L{i}.self", a)
            cx, p_cross = orig_attn(layer.crossattention, a, ctx, B, Lx, Lc, c_lens, fc)
            tap(f"{tag}.L{i}.cross", cx)
            x = orig_ffn(layer, cx, fc)
            tap(f"{tag}.L{i}.ffn", x)
            attns.append((p_self, p_cross))
        return x, attns

    M._cross_encoder = cross
    M.ops.enable_branch_streams(False)
    try:
        o = prod(b, task, True)
        for k in ("txt_embeds", "pano_embeds", "pano_fused_embeds", "gmap_embeds", "vp_embeds", "logits",
                  "global_logits", "local_logits", "fused_logits"):
            if k in o:
                tap("out." + k, o[k])
        o["loss"].float().mean().backward()
    finally:
        M._cross_encoder = orig_cross
    torch.cuda.synchronize()
    pg = {n: p.grad.detach().float().clone() for n, p in prod.named_parameters() if p.grad is not None}
    return acts, grads, pg


def rel(a, b):
    fin = torch.isfinite(b)
    a, b = a[fin], b[fin]
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


a32, g32, p32 = run(prod32)
a16, g16, p16 = run(prod16)
print(f"== task {task}: relative error bf16 vs fp32 (same product code) ==")
print(f"{'tensor':28s} {'activation':>11s} {'gradient':>11s}   |grad32|")
for n in a32:
    ga = rel(a16[n], a32[n]) if n in a16 else float('nan')
    gg = rel(g16[n].reshape(g32[n].shape), g32[n]) if (n in g16 and n in g32) else float('nan')
    print(f"{n:28s} {ga:11.4f} {gg:11.4f}   {g32[n].norm().item() if n in g32 else 0:.3e}")
rows = sorted(((rel(p16[n], p32[n]), n) for n in p32 if p32[n].norm() > 1e-9), reverse=True)
print("worst parameter gradients:")
for r, n in rows[:12]:
    print(f"  {r:.4f}  {n}")
print("median", rows[len(rows) // 2][0])
