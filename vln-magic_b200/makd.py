"""MAKD aggregation for pretraining: the composition of map_nav_src/r2r/agent.py:546-719
(`compute_kd_losses`, role 't2s') with the `mean` reductions of pretrain_src/optim/kd_loss.py, MKRW ability
weights (agent.py:866-869) and MKTD per-sample weights (agent.py:1013-1020).  Every MSE term of a step
(5 hidden-state pairs + all attention-map pairs) goes through ONE fused kernel launch, the logit KL through
a second, and the alpha mix (agent.py:1119) through a third."""
import torch

from . import ops
from .kd_loss import exponential_decay, invert_normalized_losses

KDL_DEFAULT = dict(kd_alpha=0.5, t_kd_alpha=0.5, kd_temperature=2.0, rw_temp=4.0,
                   t_sample_preprocess="exp", t_sample_preprocess_exp_decay=0.7, teacher_sample_hard_mining=True,
                   kdl_adaptive_ability_weight=True, kdl_adaptive_ability_weight_type="RW",
                   kdl_tasks=("txt", "img", "local", "global", "predict"), kdl_task_types=("emb", "attn"))

NAMES = ("txt_emb_loss", "txt_attn_loss", "img_emb_loss", "avg_img_emb_loss", "img_attn_loss", "global_emb_loss",
         "global_attn_loss", "local_emb_loss", "local_attn_loss", "predict_loss")
LEARNED_WEIGHTS = ("kdl_txt_weight", "kdl_img_weight", "kdl_global_weight", "kdl_local_weight", "kdl_predict_weight")


def kdl_config(kdl=None):
    k = dict(KDL_DEFAULT)
    if kdl:
        k.update({kk: kdl[kk] for kk in kdl.keys()} if hasattr(kdl, "keys") else vars(kdl))
    return k


def mkrw_weights(rw_temp=4.0, device="cuda", generator=None, out=None, ring=None):
    """softmax(randn(5)/rw_temp)*5, order [txt, img, global, local, predict] (agent.py:866-869); drawn per step
    per rank like the reference.  Without `out`: python floats (they scale kernel arguments).  With `out` (a
    device tensor [5]): drawn on the host generator, written into `out` in place and `out` is returned, so
    kernels captured in a CUDA graph read the current step's weights."""
    r = torch.randn(5, generator=generator)
    w = torch.softmax(r / rw_temp, 0) * 5
    if out is None:
        return w.tolist()
    if ring is not None:
        return ring.upload(w, out)  # pinned ring: no host sync, safe when the host runs ahead of the GPU
    out.copy_(w)
    return out


def grad_weights(current_iter_grads, rw_temp=4.0):
    """The 'grad' flavour of the adaptive ability weights (agent.py:856-863): softmax(-grads / rw_temp) * 5 over the
    per-ability gradient magnitudes of the previous iteration, key order [txt, img, local, global, action] as in the
    reference (:858) -- the result is used positionally like the RW draw."""
    g = torch.tensor([-float(current_iter_grads[key]) for key in ("txt", "img", "local", "global", "action")])
    return (torch.softmax(g / rw_temp, 0) * 5).tolist()


def mktd_weights(t_sample_loss, decay=0.7, preprocess="exp"):
    """MKTD per-sample weights (agent_base.py:172-175 selects the function, agent.py:1009,1019 call it with
    `decay_rate=`): 'exp' -> exponential_decay, 'norm' -> invert_normalized_losses (which ignores decay_rate)."""
    if preprocess == "exp":
        return exponential_decay(t_sample_loss, decay_rate=decay)
    if preprocess == "norm":
        return invert_normalized_losses(t_sample_loss, decay_rate=decay)
    raise ValueError("t_sample_preprocess must be 'exp' or 'norm' (map_nav_src/r2r/parser.py:169)")


def _sample_weights(out, k):
    if not k["teacher_sample_hard_mining"]:
        return None
    return mktd_weights(out["sample_loss"], k["t_sample_preprocess_exp_decay"], k.get("t_sample_preprocess", "exp"))


def _w_for(x, w, strict=False):
    if w is None or x.shape[0] == w.shape[0]:
        return w
    if strict:  # map_nav_src/utils/kd_loss.py:16-17
        raise ValueError("Shape mismatch between sample weights and inputs")
    return None  # pretrain_src/optim/kd_loss.py:15-16: silently unweighted


class _Softplus5(torch.autograd.Function):
    """w = softplus(p) for the five learned ability weights (agent.py:585,618,678,681,713).  Five scalars of glue on
    the non-default 'learned_weight' branch: plain tensor arithmetic, no kernel of ours."""

    @staticmethod
    def forward(ctx, *ps):
        p = torch.cat([x.detach().reshape(1).float() for x in ps])
        ctx.save_for_backward(p)
        return torch.where(p > 20, p, torch.log1p(torch.exp(p)))

    @staticmethod
    def backward(ctx, dw):
        (p,) = ctx.saved_tensors
        g = dw * torch.sigmoid(p)
        return tuple(g[i:i + 1] for i in range(p.numel()))


def ability_weights(student, k, rw, weight_owner=None):
    """-> (host multipliers [5], device multipliers [5] or None, divisor of the two image embedding losses).
    `weight_owner`: the model whose learned kdl_*_weight parameters are used -- the LEARNER (`s_model`, agent.py:553,
    557): the small model in role t2s (default), the large model in role s2t.
    agent.py:583-593, 616-625, 675-693, 710-717: 'RW' / 'grad' multiply by softmax_weights[a] (image embedding losses
    NOT halved); 'learned_weight' multiplies by softplus(s_model.kdl_<a>_weight) and halves the two image embedding
    losses; without kdl_adaptive_ability_weight every weight is 1 and the two image embedding losses are halved."""
    if not k.get("kdl_adaptive_ability_weight", True):
        return [1.0] * 5, None, 2.0
    kind = k.get("kdl_adaptive_ability_weight_type", "RW")
    if kind in ("RW", "grad"):
        if rw is None:
            raise ValueError("adaptive ability weights of type %s need the step's weights (mkrw_weights / grad_weights)"
                             % kind)
        if torch.is_tensor(rw):
            return [1.0] * 5, rw.float().contiguous(), 1.0
        return [float(x) for x in rw], None, 1.0
    if kind == "learned_weight":
        owner = weight_owner if weight_owner is not None else student
        ps = [getattr(owner.bert, n, None) for n in LEARNED_WEIGHTS]
        if any(p is None for p in ps):
            raise ValueError("kdl_adaptive_ability_weight_type='learned_weight' needs the learner's kdl_*_weight "
                             "parameters (build the model with that kdl config)")
        return [1.0] * 5, _Softplus5.apply(*[p.reshape(1) for p in ps]), 2.0
    raise ValueError("kdl_adaptive_ability_weight_type must be RW, grad or learned_weight")


def compute_kd_losses(student, s_out, t_out, task, rw, t_w, kdl=None, role="t2s", weight_owner=None):
    """-> (named dict of per-ability scalars as a [10] tensor view, mse_total scalar, kl scalar or None).
    `rw`: the 5 MKRW ability weights, either python floats (baked into the kernel arguments) or a DEVICE tensor
    [5] (read by the kernels at run time, so a captured CUDA graph follows the per-step draw).
    `student` is always the SMALL model (it owns the up-projections).  role 't2s' (agent.py:550-552): the small
    model learns, prediction = proj(s_out), target = t_out.detach().  role 's2t' (agent.py:553-556, ICoD): the
    LARGE model learns; pass (s_out, t_out) = (large model's outputs, small model's outputs) as agent.py:1022
    does; prediction = s_out, target = proj(t_out).detach() (agent.py:571,605-606,647,665) and `t_w` are the small
    model's MKTD weights.
    Batches padded by graph_index.pad_batch carry `pano_row_scale` / `row_scale` (model outputs): padded panoramas
    and padded masked-token rows get weight 0 and the means are taken over the real rows.
    Reductions (`kdl['kd_loss_type']`): None (default) = the pretraining file's `mean` with its silent fallback on a
    weight / batch mismatch (pretrain_src/optim/kd_loss.py); 'mean' / 'sum' = the fine-tune file's, which raises on
    the mismatch (map_nav_src/utils/kd_loss.py; agent.py:554 for role t2s, always 'mean' for role s2t, :557).
    Pinned against the reference's own compute_kd_losses source: tests/test_makd_agent_pinned.py."""
    k = kdl_config(kdl)
    aw, aw_dev, img_div = ability_weights(student, k, rw, weight_owner)
    lt = k.get("kd_loss_type")
    if lt not in (None, "mean", "sum"):
        raise ValueError("Unsupported loss_type. Choose 'sum' or 'mean'.")
    if role == "s2t" and lt is not None:
        lt = "mean"
    total = lt == "sum"

    def sdev(i):
        return aw_dev.detach()[i:i + 1] if aw_dev is not None else None

    bert = student.bert
    emb, att = "emb" in k["kdl_task_types"], "attn" in k["kdl_task_types"]
    tasks = k["kdl_tasks"]
    pairs, owner, ability = [], [], []
    pano_scale = s_out.get("pano_row_scale")

    def row_w(x, pano):
        w = _w_for(x, t_w, strict=lt is not None)
        if pano and pano_scale is not None and pano_scale.shape[0] == x.shape[0]:
            return ops.row_weights(w, None, pano_scale, x.shape[0])
        return w

    def add_emb(name, proj, s, t, ri, div=1.0, pano=False):
        if not emb:
            return
        if role == "t2s":
            ps = ops.linear(s, proj.weight, proj.bias)
            t = t.detach()
        else:
            with torch.no_grad():
                ps, t = s, ops.linear(t.detach(), proj.weight, proj.bias)
        pairs.append((ps, t, row_w(ps, pano), aw[ri] / ((1 if total else ps.numel()) * div), sdev(ri)))
        owner.append(name)
        ability.append(ri)

    def add_attn(name, s_list, t_list, ri, n_layers, pano=False):
        if not att or not s_list:
            return
        items = []
        for sl, tl in zip(s_list[:n_layers], t_list[:n_layers]):
            if isinstance(sl, tuple):
                items += [(sl[0], tl[0]), (sl[1], tl[1])]
            else:
                items.append((sl, tl))
        numel = 1 if total else sum(a.numel() for a, _ in items)
        w = row_w(items[0][0], pano) if items else None
        for a, b in items:
            pairs.append((a, b.detach(), w, aw[ri] / numel, sdev(ri)))
            owner.append(name)
            ability.append(ri)

    # agent.py:560 -- the attention maps are compared on their first min(layers) layers
    min_len = min(len(s_out["txt_attn_list"]), len(t_out["txt_attn_list"])) if att else 0
    if "txt" in tasks:
        add_emb("txt_emb_loss", bert.txt_emb_w, s_out["txt_embeds"], t_out["txt_embeds"], 0)
        add_attn("txt_attn_loss", s_out["txt_attn_list"], t_out["txt_attn_list"], 0, min_len)
    if "img" in tasks:
        add_emb("img_emb_loss", bert.kdl_img_w, s_out["pano_embeds"], t_out["pano_embeds"], 1, img_div, pano=True)
        add_emb("avg_img_emb_loss", bert.kdl_avg_img_w, s_out["pano_fused_embeds"], t_out["pano_fused_embeds"], 1,
                img_div, pano=True)
        if att and len(s_out["img_attn_list"]) != len(t_out["img_attn_list"]):
            raise ValueError("img_attns of teacher and student must have the same shape (agent.py:628)")
        add_attn("img_attn_loss", s_out["img_attn_list"], t_out["img_attn_list"], 1, len(s_out["img_attn_list"]),
                 pano=True)
    mlm = task.startswith("mlm")
    gw, lw = (bert.gmap_txt_w, bert.vp_txt_w) if mlm else (bert.global_cross_w, bert.local_cross_w)
    nx = min(len(s_out["gmap_attn_list"]), len(t_out["gmap_attn_list"]), max(min_len, 0)) if att else 0
    if "global" in tasks:
        add_emb("global_emb_loss", gw, s_out["gmap_embeds"], t_out["gmap_embeds"], 2)
        add_attn("global_attn_loss", s_out["gmap_attn_list"], t_out["gmap_attn_list"], 2, nx)
    if "local" in tasks:
        add_emb("local_emb_loss", lw, s_out["vp_embeds"], t_out["vp_embeds"], 3)
        add_attn("local_attn_loss", s_out["vp_attn_list"], t_out["vp_attn_list"], 3, nx)
    per_seg, mse_total = ops.makd_mse(pairs, aw_dev, ability) if pairs else (None, None)
    kl = None
    if "predict" in tasks:
        s_log, t_log = s_out["logits"], t_out["logits"].detach()
        R, C = s_log.shape
        w = t_w
        if mlm:
            # per-row weights for the [n_masked, vocab] logits: the row's sample weight x the padding mask
            rs = s_out.get("row_scale")
            if t_w is not None or rs is not None:
                w = ops.row_weights(t_w, s_out["row_sample"] if t_w is not None else None, rs, R)
        T = float(k["kd_temperature"])
        scale = (T * T if total else T * T / R if t_w is not None else T * T / (R * C)) * aw[4]
        kl = ops.makd_kl(s_log, t_log, T, w, scale, sdev(4), aw_dev)
    return dict(per_seg=per_seg, owner=owner, mse_total=mse_total, kl=kl)


def named_losses(res):
    """Host-side (logging / tests): fold the per-segment vector into the reference's 10 named scalars."""
    out = {n: 0.0 for n in NAMES}
    if res["per_seg"] is not None:
        vals = res["per_seg"].detach().float().cpu().tolist()
        for name, v in zip(res["owner"], vals):
            out[name] += v
    if res["kl"] is not None:
        out["predict_loss"] = float(res["kl"].detach())
    return out


def distill_step_loss(student, teacher, batch, task, rw, kdl=None):
    """Forward of the (missing upstream) distillation step, SURVEY.md 3.2.
    Returns (mix [total, sup_mean, kd_total], kd result dict, s_out, t_out)."""
    k = kdl_config(kdl)

    def t_fwd():
        with torch.no_grad():
            return teacher(batch, task, True, output_kd=True)

    # the frozen teacher's forward and the student's forward are independent until the losses: two stream branches
    # (the student's small latency-bound kernels fill the gaps of the teacher's wide GEMMs)
    t_out, s_out = ops.run_branches(t_fwd, lambda: student(batch, task, True, output_kd=True))
    t_w = _sample_weights(t_out, k)
    res = compute_kd_losses(student, s_out, t_out, task, rw, t_w, k)
    mix = ops.loss_mix(res["mse_total"], res["kl"], s_out["loss"], k["kd_alpha"], s_out.get("loss_inv_n"))
    return mix, res, s_out, t_out


def teacher_forward(teacher, batch, task):
    """The frozen teacher's half of the distillation step (no grad, KD outputs on)."""
    with torch.no_grad():
        return teacher(batch, task, True, output_kd=True)


def student_distill_loss(student, t_out, batch, task, rw, kdl=None):
    """The student's half: forward, MAKD losses against the given teacher outputs, alpha mix.
    `distill_step_loss` == `student_distill_loss(student, teacher_forward(teacher, batch, task), ...)`; the split
    lets the stepper run the frozen teacher's forward of the NEXT batch while this batch back-propagates.
    Returns (mix [total, sup_mean, kd_total], kd result dict, s_out)."""
    k = kdl_config(kdl)
    s_out = student(batch, task, True, output_kd=True)
    t_w = _sample_weights(t_out, k)
    res = compute_kd_losses(student, s_out, t_out, task, rw, t_w, k)
    mix = ops.loss_mix(res["mse_total"], res["kl"], s_out["loss"], k["kd_alpha"], s_out.get("loss_inv_n"))
    return mix, res, s_out


def icod_step_loss(student, teacher, batch, task, rw, t_rw, kdl=None):
    """ICoD co-update (`--train_kdl_teacher`; agent.py:1019-1022, 1136-1149, agent_base.py:260-279): the large
    model's forward runs WITH grad and one step produces two losses, one per model, from two disjoint autograd
    graphs (every cross-model target is detached).  Returns (mix_s, mix_t, res_s, res_t, s_out, t_out); each mix is
    the device vector [total, supervised_mean, kd_total] of `ops.loss_mix`."""
    k = kdl_config(kdl)
    t_out = teacher(batch, task, True, output_kd=True)
    s_out = student(batch, task, True, output_kd=True)
    t_w, s_w = _sample_weights(t_out, k), _sample_weights(s_out, k)
    res_s = compute_kd_losses(student, s_out, t_out, task, rw, t_w, k, role="t2s")
    # agent.py:553,557: in role s2t the learned ability weights are the LARGE model's own kdl_*_weight parameters; a
    # large model built without them (pretraining teacher configs carry no kdl block) uses the small model's
    owner = teacher if getattr(teacher.bert, LEARNED_WEIGHTS[0], None) is not None else None
    res_t = compute_kd_losses(student, t_out, s_out, task, t_rw, s_w, k, role="s2t", weight_owner=owner)
    mix_s = ops.loss_mix(res_s["mse_total"], res_s["kl"], s_out["loss"], k["kd_alpha"], s_out.get("loss_inv_n"))
    mix_t = ops.loss_mix(res_t["mse_total"], res_t["kl"], t_out["loss"], k["t_kd_alpha"], t_out.get("loss_inv_n"))
    return mix_s, mix_t, res_s, res_t, s_out, t_out
