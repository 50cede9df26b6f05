"""Golden vectors of the FROZEN oracle (oracle/magic_oracle.py) on seeded synthetic inputs and weights.
The reference cannot produce these (its model files are absent, readme.md:75), so the golden file pins OUR
restatement against drift; the reference-pinned parts (KD arithmetic, collate) have their own golden file.
Run:  python tests/golden/gen_model_golden.py     (CPU, ~20 s; output committed as model_golden.pt)"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import magic_b200  # noqa: E402,F401
from magic_b200 import synth  # noqa: E402
from oracle import magic_oracle as O  # noqa: E402

RW = [1.3, 0.6, 1.1, 0.9, 1.1]


def build(device="cpu"):
    cfg_t = O.make_config(256, role="teacher", hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    cfg_s = O.make_config(128, role="student", teacher_hidden_size=256, hidden_dropout_prob=0.0,
                          attention_probs_dropout_prob=0.0)
    torch.manual_seed(0)
    teacher = O.GlocalTextPathCMTPreTraining(cfg_t).to(device).eval()
    torch.manual_seed(1)
    student = O.GlocalTextPathCMTPreTraining(cfg_s).to(device).train()
    return cfg_t, cfg_s, teacher, student


def run(task, teacher, student, device="cpu"):
    b = synth.make_batch(task, 4, seed=2026)
    b = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in b.items()}
    total, sup, kd, L, s_out, t_out = O.distill_step_loss(student, teacher, b, task, torch.tensor(RW, device=device))
    rec = dict(total=total.detach().cpu(), sup=sup.detach().cpu(), kd=kd.detach().cpu(),
               named={k: float(v) for k, v in L.items()}, loss=s_out["loss"].detach().cpu(),
               t_sample_loss=t_out["sample_loss"].detach().cpu())
    if task == "sap":
        for k in ("global_logits", "local_logits", "fused_logits"):
            rec[k] = s_out[k].detach().cpu().clone()
    else:
        rec["logits_head"] = s_out["logits"][:, :64].detach().cpu().clone()
        rec["logits_lse"] = torch.logsumexp(s_out["logits"], 1).detach().cpu()
        rec["logits_argmax"] = s_out["logits"].argmax(1).cpu()
    return rec


def main():
    _, _, teacher, student = build()
    out = {t: run(t, teacher, student) for t in ("sap", "mlm")}
    out["rw"] = RW
    torch.save(out, os.path.join(HERE, "model_golden.pt"))
    print({t: float(out[t]["total"]) for t in ("sap", "mlm")})


if __name__ == "__main__":
    main()
