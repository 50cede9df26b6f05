"""The training loop `pretrain_src/train_r2r_magic.py` sets up but does not ship (its body is missing after
:399-401), as a library: per-step LR schedule, the distillation step (train_step.PretrainStepper),
the meters the reference declares and never feeds (:380-390), periodic validation with the reference's own metric
definitions (:412-587) and checkpointing with the reference's file layout (utils/save.py:23-74).

  RunningMeter            pretrain_src/utils/logger.py:66-94 (EMA 0.99, NaN dropped)
  ModelSaver              pretrain_src/utils/save.py:23-74   (`module.` stripped, CPU tensors, train_state_*.pt)
  validate / validate_*   pretrain_src/train_r2r_magic.py:412-587 (same keys: loss/acc/tok_per_s, gloss..facc)
  train                   the missing loop: task schedule -> prefetch -> step -> meters -> validate -> save

Validation losses are computed by the CUDA kernels of this package (ops.cross_entropy / soft_cross_entropy); only the
argmax / comparison bookkeeping of the metrics uses torch tensor methods."""
import json
import math
import os
import time
from collections import defaultdict

import torch
import torch.distributed as dist

from . import ops
from .optim import get_lr_sched
from .parallel import task_schedule


class RunningMeter(object):
    """Running mean of a scalar (utils/logger.py:66-94): val <- value*(1-s) + val*s, NaN updates are dropped."""

    def __init__(self, name, val=None, smooth=0.99):
        self._name, self._sm, self._val = name, smooth, val

    def __call__(self, value):
        val = value if self._val is None else value * (1 - self._sm) + self._val * self._sm
        if not math.isnan(val):
            self._val = val

    def __str__(self):
        return f"{self._name}: {self.val:.4f}"

    @property
    def val(self):
        return 0 if self._val is None else self._val

    @property
    def name(self):
        return self._name


class ModelSaver(object):
    """utils/save.py:23-74: `<prefix>_<step>.pt`, `<prefix>_latest.pt`, `<prefix>_best.pt` hold the model's
    state_dict with a DDP `module.` prefix stripped and tensors on the CPU; `train_state_*.pt` = {step, optimizer}."""

    def __init__(self, output_dir, prefix="model_step", suffix="pt"):
        self.output_dir, self.prefix, self.suffix = output_dir, prefix, suffix
        os.makedirs(output_dir, exist_ok=True)

    @staticmethod
    def _state(model):
        out = {}
        for k, v in model.state_dict().items():
            k = k[7:] if k.startswith("module.") else k
            out[k] = v.detach().cpu().clone() if isinstance(v, torch.Tensor) else v
        return out

    def save(self, model, step, optimizer=None):
        torch.save(self._state(model), os.path.join(self.output_dir, f"{self.prefix}_{step}.{self.suffix}"))
        if optimizer is not None:
            torch.save({"step": step, "optimizer": optimizer.state_dict()},
                       os.path.join(self.output_dir, f"train_state_{step}.pt"))

    def save_latest(self, model, step, optimizer=None, is_max=False):
        tag = "best" if is_max else "latest"
        torch.save(self._state(model), os.path.join(self.output_dir, f"{self.prefix}_{tag}.{self.suffix}"))
        if optimizer is not None:
            name = f"train_state_best_{step}.pt" if is_max else "train_state_latest.pt"
            torch.save({"step": step, "optimizer": optimizer.state_dict()}, os.path.join(self.output_dir, name))


def save_training_meta(opts, model_config):
    """utils/save.py:12-20."""
    os.makedirs(os.path.join(opts.output_dir, "logs"), exist_ok=True)
    os.makedirs(os.path.join(opts.output_dir, "ckpts"), exist_ok=True)
    with open(os.path.join(opts.output_dir, "logs", "training_args.json"), "w") as f:
        json.dump({k: (v if isinstance(v, (int, float, str, bool, list, dict, type(None))) else str(v))
                   for k, v in vars(opts).items()}, f, indent=4)
    with open(os.path.join(opts.output_dir, "logs", "model_config.json"), "w") as f:
        cfg = model_config if isinstance(model_config, dict) else vars(model_config)
        json.dump({k: (sorted(v) if isinstance(v, set) else v) for k, v in cfg.items()
                   if isinstance(v, (int, float, str, bool, list, dict, set, type(None)))}, f, indent=4)


def all_gather(x):
    """utils/distributed.py:95-135 for picklable python values: a list with one entry per rank."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        out = [None] * dist.get_world_size()
        dist.all_gather_object(out, x)
        return out
    return [x]


# ---------------------------------------------------------------------------------------------------
# validation (train_r2r_magic.py:412-587; same metric names and definitions)
# ---------------------------------------------------------------------------------------------------
def _ce_sum(logits, labels, ignore_index=-100):
    return float(ops.cross_entropy(logits, labels, ignore_index).sum())


@torch.no_grad()
def validate_mlm(model, val_loader):
    val_loss = n_correct = n_word = 0
    st = time.time()
    for batch in val_loader:
        scores = model(batch, task="mlm", compute_loss=False)["predict"]
        labels = batch["txt_labels"]
        labels = labels[labels != -1]
        val_loss += _ce_sum(scores, labels, -1)
        n_correct += int((scores.max(dim=-1)[1] == labels).sum())
        n_word += labels.numel()
    val_loss, n_correct, n_word = sum(all_gather(val_loss)), sum(all_gather(n_correct)), sum(all_gather(n_word))
    tot = max(time.time() - st, 1e-9)
    return {"loss": val_loss / max(n_word, 1), "acc": n_correct / max(n_word, 1), "tok_per_s": n_word / tot}


@torch.no_grad()
def validate_mrc(model, val_loader):
    val_loss = n_feat = tot_score = 0
    st = time.time()
    for batch in val_loader:
        view_logits, view_targets, _, _ = model(batch, task="mrc", compute_loss=False)
        val_loss += float(ops.soft_cross_entropy(view_logits, view_targets).sum())  # KL(targets || softmax), summed
        tot_score += int((view_logits.max(dim=-1)[1] == view_targets.max(dim=-1)[1]).sum())
        n_feat += int(batch["vp_view_mrc_masks"].sum())
    val_loss, tot_score, n_feat = sum(all_gather(val_loss)), sum(all_gather(tot_score)), sum(all_gather(n_feat))
    tot = max(time.time() - st, 1e-9)
    return {"loss": val_loss / max(n_feat, 1), "acc": tot_score / max(n_feat, 1), "feat_per_s": n_feat / tot}


@torch.no_grad()
def validate_sap(model, val_loader):
    sums = defaultdict(float)
    n_data = 0
    st = time.time()
    for batch in val_loader:
        o = model(batch, task="sap", compute_loss=False)
        gl, ll, fl, ga, la = (o[k] for k in ("global_logits", "local_logits", "fused_logits", "global_act_labels",
                                             "local_act_labels"))
        sums["gloss"] += _ce_sum(gl, ga)
        sums["lloss"] += _ce_sum(ll, la)
        sums["floss"] += _ce_sum(fl, ga)
        sums["gacc"] += int((torch.argmax(gl, 1) == ga).sum())
        sums["lacc"] += int((torch.argmax(ll, 1) == la).sum())
        sums["facc"] += int((torch.argmax(fl, 1) == ga).sum())
        n_data += len(ga)
    n_data = sum(all_gather(n_data))
    log = {k: sum(all_gather(sums[k])) / max(n_data, 1) for k in ("gloss", "lloss", "floss", "gacc", "lacc", "facc")}
    log["tok_per_s"] = n_data / max(time.time() - st, 1e-9)
    return log


@torch.no_grad()
def validate_cfp(model, val_loader, temperature):
    sums = defaultdict(float)
    n_data = 0
    st = time.time()
    for batch in val_loader:
        g, v, f, t = model(batch, task="cfp", compute_loss=False)
        tgt = torch.arange(len(g), device=g.device)
        for key, a in (("g", g), ("l", v), ("f", f)):
            sim = ops.matmul_nt(a, t, 1.0 / temperature)
            sim_t = ops.matmul_nt(t, a, 1.0 / temperature)
            sums[key + "loss"] += (_ce_sum(sim, tgt) + _ce_sum(sim_t, tgt)) / 2.0
            sums[key + "acc"] += int((torch.argmax(sim.float(), 1) == tgt).sum())
        n_data += len(tgt)
    n_data = sum(all_gather(n_data))
    log = {k: sum(all_gather(sums[k])) / max(n_data, 1) for k in ("gloss", "lloss", "floss", "gacc", "lacc", "facc")}
    log["tok_per_s"] = n_data / max(time.time() - st, 1e-9)
    return log


def validate(model, val_dataloaders, setname="", max_metrix=None, tem=None, log=None):
    """train_r2r_magic.py:412-437.  Returns {f'val{setname}_{task}_{k}': v} (and, like the reference, the updated
    best unseen SAP accuracy when `max_metrix` is given)."""
    was_training = model.training
    model.eval()
    out, flag = {}, False
    for task, loader in val_dataloaders.items():
        if task.startswith("mlm"):
            val_log = validate_mlm(model, loader)
        elif task.startswith("mrc"):
            val_log = validate_mrc(model, loader)
        elif task.startswith("sap"):
            val_log = validate_sap(model, loader)
            if setname == "_unseen" and max_metrix is not None and val_log["facc"] >= max_metrix:
                max_metrix, flag = val_log["facc"], True
        elif task.startswith("cfp"):
            val_log = validate_cfp(model, loader, tem if tem is not None else model.config.cfp_temperature)
        else:
            raise ValueError(f"Undefined task {task}")
        out.update({f"val{setname}_{task}_{k}": v for k, v in val_log.items()})
    if was_training:
        model.train()
    if log is not None:
        log(out)
    return (out, max_metrix, flag) if max_metrix is not None else out


# ---------------------------------------------------------------------------------------------------
# the loop
# ---------------------------------------------------------------------------------------------------
class MetaLoader:
    """data/loader.py:18-75: samples the next task from `mix_ratio` and yields (task, batch).  The reference
    broadcasts the rank-0 draw every step; here every rank derives the same seeded schedule (parallel.task_schedule),
    so no collective is needed.  `loaders`: {task: iterable of host batches (graph_index.prepare_batch output)} or,
    like the reference (:29-33), {task: (iterable, ratio, pre_epoch)} with `pre_epoch(epoch_id)` called before a
    task's loader restarts (the sampler's `set_epoch`, :66-71).  The task is re-drawn every `accum_steps` steps
    (:55-56).  tests/test_host_pinned_live.py replays the schedule through the reference's own class."""

    def __init__(self, loaders, mix_ratio=None, seed=0, num_steps=1 << 20, accum_steps=1):
        self.tasks = list(loaders.keys())
        self.loaders, self.pre_epoch, ratios = {}, {}, []
        for t, l in loaders.items():
            l, r, p = l if isinstance(l, tuple) else (l, 1, None)
            self.loaders[t], self.pre_epoch[t] = l, p
            ratios.append(r)
        self.iters = {t: iter(l) for t, l in self.loaders.items()}
        if mix_ratio is not None:
            ratios = list(mix_ratio)
        self.accum_steps = max(int(accum_steps), 1)
        self.num_steps = num_steps
        self.schedule = task_schedule(seed, (num_steps + self.accum_steps - 1) // self.accum_steps, self.tasks, ratios)
        self.step = 0
        self.epoch_id = 0

    def __iter__(self):
        while self.step < self.num_steps:
            task = self.schedule[self.step // self.accum_steps]
            self.step += 1
            try:
                batch = next(self.iters[task])
            except StopIteration:  # a new epoch of that task
                self.epoch_id += 1
                if self.pre_epoch[task] is not None:
                    self.pre_epoch[task](self.epoch_id)
                self.iters[task] = iter(self.loaders[task])
                batch = next(self.iters[task])
            yield task, batch


def train(opts, stepper, meta_loader, val_dataloaders=None, val2_dataloaders=None, model_saver=None, log=print,
          start_step=0):
    """The step loop.  `opts`: learning_rate, warmup_steps, num_train_steps, log_steps, valid_steps (the reference's
    names, config/r2r_magic_pretrain.json:8-24).  One-deep host->device prefetch like the reference's PrefetchLoader
    (data/loader.py:90-124); the next batch is also announced to the stepper so a frozen teacher's forward of step
    i+1 runs under step i's backward.  Returns the meters."""
    student = stepper.student
    kdl_tasks = list(stepper.kdl["kdl_tasks"]) if stepper.teacher is not None else []
    task2loss = {}
    for task in meta_loader.tasks:
        task2loss[task] = {k: RunningMeter(f"loss/{task}/{k}") for k in kdl_tasks}
        if stepper.teacher is not None:
            task2loss[task]["kdl_loss"] = RunningMeter(f"loss/{task}/kdl_loss")
        task2loss[task]["supervised_loss"] = RunningMeter(f"loss/{task}/supervised_loss")
        task2loss[task]["total_loss"] = RunningMeter(f"loss/{task}/total_loss")
    n_examples = defaultdict(int)
    global_step = start_step
    max_unseen_facc, max_unseen_iter = 0.0, 0
    pending = []  # (task, device loss vector): read back at log time, so the loop never drains the GPU per step
    start = time.time()
    it = iter(meta_loader)

    def fetch():
        try:
            task, hb = next(it)
        except StopIteration:
            return None
        return task, stepper.prefetch(task, hb)

    cur = fetch()
    while cur is not None and global_step < opts.num_train_steps:
        nxt = fetch()
        task, handle = cur
        lr = get_lr_sched(global_step, opts)
        out = stepper.step(task, handle, lr=lr, next=nxt)
        n_examples[task] += int(handle.batch["txt_ids"].shape[0])
        pending.append((task, out.clone()))
        global_step += 1
        if global_step % opts.log_steps == 0 or global_step == opts.num_train_steps:
            for t, o in pending:
                vals = o.tolist()
                task2loss[t]["total_loss"](vals[0])
                task2loss[t]["supervised_loss"](vals[1])
                if "kdl_loss" in task2loss[t]:
                    task2loss[t]["kdl_loss"](vals[2])
            pending.clear()
            named = stepper.last_named_losses()
            if named is not None:
                t_last = task
                group = {"txt": ("txt_emb_loss", "txt_attn_loss"), "img": ("img_emb_loss", "avg_img_emb_loss", "img_attn_loss"),
                         "global": ("global_emb_loss", "global_attn_loss"), "local": ("local_emb_loss", "local_attn_loss"),
                         "predict": ("predict_loss",)}
                for k in kdl_tasks:
                    task2loss[t_last][k](sum(named[n] for n in group.get(k, ())))
            ex_per_sec = sum(n_examples.values()) / max(time.time() - start, 1e-9)
            log({"step": global_step, "lr": lr, "ex_per_s": ex_per_sec, "grad_norm": stepper.opt.grad_norm(),
                 **{m.name: m.val for t in task2loss.values() for m in t.values()}})
        if getattr(opts, "valid_steps", 0) and global_step % opts.valid_steps == 0:
            if val_dataloaders:
                validate(student, val_dataloaders, setname="_seen", log=log)
            if val2_dataloaders:
                _, max_new, flag = validate(student, val2_dataloaders, setname="_unseen", max_metrix=max_unseen_facc,
                                            log=log)
                if flag:
                    max_unseen_facc, max_unseen_iter = max_new, global_step
                    if model_saver is not None:
                        model_saver.save_latest(student, global_step, stepper.opt, is_max=True)
            if model_saver is not None:
                model_saver.save_latest(student, global_step, stepper.opt)
        cur = nxt
    torch.cuda.synchronize()
    return dict(meters=task2loss, global_step=global_step, best_unseen_facc=max_unseen_facc,
                best_unseen_iter=max_unseen_iter)
