"""Generates tests/golden/kd_loss_ref.pt by IMPORTING THE REFERENCE's own KD loss files in the build container:
  /root/reference/pretrain_src/optim/kd_loss.py      (pretraining variant, mean reductions)
  /root/reference/map_nav_src/utils/kd_loss.py       (fine-tune variant, loss_type sum/mean, raises on mismatch)
Run:  python tests/golden/gen_kd_golden.py      (needs /root/reference; the output is committed)"""
import importlib.util
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def cases():
    g = torch.Generator().manual_seed(20261017)
    B = 6
    out = []
    s, t = torch.randn(B, 7, 16, generator=g), torch.randn(B, 7, 16, generator=g)
    w = torch.rand(B, generator=g)
    out.append(("mse", dict(s=s, t=t, w=None)))
    out.append(("mse", dict(s=s, t=t, w=w)))
    out.append(("mse", dict(s=s, t=t, w=w[:4])))  # mismatch: silently unweighted (pretrain) / raises (fine-tune)
    sl, tl = torch.randn(B, 20, generator=g) * 3, torch.randn(B, 20, generator=g) * 3
    sl[:, 5:9] = float("-inf")
    tl[:, 5:9] = float("-inf")
    tl[1, 12] = float("-inf")
    for T in (1.0, 2.0, 4.0):
        out.append(("kd", dict(s=sl, t=tl, T=T, w=None)))
        out.append(("kd", dict(s=sl, t=tl, T=T, w=w)))
    out.append(("exp", dict(x=torch.rand(B, generator=g) * 5, rate=0.7)))
    out.append(("inv", dict(x=torch.rand(B, generator=g) * 5)))
    return out


def main():
    pre = load("/root/reference/pretrain_src/optim/kd_loss.py", "ref_kd_pre")
    fin = load("/root/reference/map_nav_src/utils/kd_loss.py", "ref_kd_fin")
    rec = []
    for kind, a in cases():
        r = dict(kind=kind, args=a)
        if kind == "mse":
            r["pretrain"] = pre.mse_loss(a["s"], a["t"], a["w"])
            for lt in ("sum", "mean"):
                try:
                    r["finetune_" + lt] = fin.mse_loss(a["s"], a["t"], a["w"], loss_type=lt)
                except ValueError as e:
                    r["finetune_" + lt] = "ValueError"
        elif kind == "kd":
            r["pretrain"] = pre.kd_loss(a["s"], a["t"], temperature=a["T"], t_sample_weights=a["w"])
            for lt in ("sum", "mean"):
                r["finetune_" + lt] = fin.kd_loss(a["s"], a["t"], temperature=a["T"], t_sample_weights=a["w"], loss_type=lt)
        elif kind == "exp":
            r["pretrain"] = pre.exponential_decay(a["x"], a["rate"])
        else:
            r["pretrain"] = pre.invert_normalized_losses(a["x"])
        rec.append(r)
    torch.save(rec, os.path.join(HERE, "kd_loss_ref.pt"))
    print("wrote", len(rec), "cases")


if __name__ == "__main__":
    main()
