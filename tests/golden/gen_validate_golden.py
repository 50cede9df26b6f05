"""Generates tests/golden/validate_ref.pt by EXECUTING THE REFERENCE's own validation functions in the build container:
the SOURCE of `validate_mlm`, `compute_accuracy_for_soft_targets`, `validate_mrc`, `validate_sap`, `validate_cfp`
(/root/reference/pretrain_src/train_r2r_magic.py:440-577; the module itself cannot be imported: easydict, tensorboardX
and the absent model package) runs on a stub model that replays canned `compute_loss=False` outputs -- two batches per
task, -inf-masked SAP logits, an ignored (-100) SAP label, -1 MLM labels, soft MRC targets.

Run:  python tests/golden/gen_validate_golden.py      (needs /root/reference; the output is committed)"""
import os
import re
import time
import types

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/pretrain_src/train_r2r_magic.py"
OUT = os.path.join(HERE, "validate_ref.pt")
FUNCS = ("validate_mlm", "compute_accuracy_for_soft_targets", "validate_mrc", "validate_sap", "validate_cfp")
TEMPERATURE = 0.07


def reference_functions():
    src = open(REF).read()
    ns = dict(torch=torch, F=F, time=time, all_gather=lambda x: [x],
              LOGGER=types.SimpleNamespace(info=lambda *a, **k: None))
    for name in FUNCS:
        m = re.search(rf"^(@torch\.no_grad\(\)\n)?def {name}\(.*?(?=^@torch\.no_grad|^def )", src, re.S | re.M)
        exec(m.group(0), ns)
    return {n: ns[n] for n in FUNCS}


def canned():
    g = torch.Generator().manual_seed(20261019)

    def r(*s):
        return torch.randn(*s, generator=g)

    mlm, mrc, sap, cfp = [], [], [], []
    for B in (5, 3):
        lab = torch.randint(0, 64, (B, 12), generator=g)
        lab[torch.rand(B, 12, generator=g) > 0.25] = -1
        lab[0, 0] = 7
        n = int((lab != -1).sum())
        pred = r(n, 64) * 2
        pred[0, 7] = 50.0  # at least one correct prediction
        mlm.append((dict(txt_labels=lab), dict(predict=pred)))

        m = torch.rand(B, 36, generator=g) < 0.15
        m[0, 0] = True
        nv = int(m.sum())
        tgt = torch.softmax(r(nv, 32) * 2, -1)
        logit = r(nv, 32)
        logit[0] = tgt[0] * 10
        mrc.append((dict(vp_view_mrc_masks=m), (logit, tgt, None, None)))

        G, Vp = 9, 8
        gl, ll = r(B, G) * 2, r(B, Vp) * 2
        gl[:, 6:] = float("-inf")
        gl[:, 2] = float("-inf")
        ll[:, 5:] = float("-inf")
        fl = gl + 0.3 * r(B, G)
        ga, la = torch.randint(3, 6, (B,), generator=g), torch.randint(0, 5, (B,), generator=g)
        ga[0] = int(gl[0].argmax())
        if B == 5:
            ga[1] = -100   # an ignored label: it still counts in n_data (train_r2r_magic.py:518)
            la[1] = -100
        sap.append((dict(), dict(global_logits=gl, local_logits=ll, fused_logits=fl, global_act_labels=ga,
                                 local_act_labels=la)))

        outs = [F.normalize(r(B, 16), dim=-1) for _ in range(4)]
        outs[0] = F.normalize(outs[3] + 0.2 * outs[0], dim=-1)  # the graph side mostly retrieves its own text
        cfp.append((dict(), tuple(outs)))
    return dict(mlm=mlm, mrc=mrc, sap=sap, cfp=cfp)


class Replay:
    """A model that replays the canned outputs in loader order."""

    def __init__(self, items, device="cpu"):
        self.items, self.i, self.device = items, 0, device

    def _to(self, o):
        if torch.is_tensor(o):
            return o.to(self.device)
        if isinstance(o, dict):
            return {k: self._to(v) for k, v in o.items()}
        if isinstance(o, tuple):
            return tuple(self._to(v) for v in o)
        return o

    def loader(self):
        return [self._to(b) for b, _ in self.items]

    def __call__(self, batch, task=None, compute_loss=True):
        assert compute_loss is False
        out = self._to(self.items[self.i][1])
        self.i += 1
        return out


def run_reference(data):
    fns = reference_functions()
    cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self  # validate_cfp builds its targets with .cuda()
    try:
        out = {}
        for task in ("mlm", "mrc", "sap"):
            m = Replay(data[task])
            out[task] = fns[f"validate_{task}"](m, m.loader())
        m = Replay(data["cfp"])
        out["cfp"] = fns["validate_cfp"](m, m.loader(), TEMPERATURE)
    finally:
        torch.Tensor.cuda = cuda
    return {t: {k: float(v) for k, v in d.items() if not k.endswith("_per_s")} for t, d in out.items()}


def main():
    data = canned()
    ref = run_reference(data)
    torch.save(dict(data=data, ref=ref, temperature=TEMPERATURE), OUT)
    print("wrote", OUT)
    for k, v in ref.items():
        print(k, v)


if __name__ == "__main__":
    main()
