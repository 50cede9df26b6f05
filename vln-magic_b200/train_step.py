"""The training step the reference's `train_r2r_magic.py` builds everything for but does not ship
(its loop body is missing after :399-401): teacher forward (no grad), student forward,
supervised task loss, MAKD losses, alpha mix, backward, gradient all-reduce, clip + AdamW.

Optionally the whole device side of a step is captured into a CUDA graph per (task, shape signature) and
replayed: at MAGIC-S sizes (h = 128) the step is launch-bound, so removing the per-kernel host cost is the
single largest win."""
import torch
import torch.distributed as dist

from . import _lib, makd, ops
from .graph_index import FLAT_KEY, INDEX_KEY, alloc_like, copy_batch_
from .optim import FusedAdamW
from .arena import ParamArena
from .parallel import FlatAllReduce, StageSync, broadcast_flat


class _Prefetched:
    """A batch whose host->device copy is in flight on the copy stream (PretrainStepper.prefetch)."""

    def __init__(self, task, batch, ready, slot, index):
        self.task, self.batch, self.ready, self.slot, self.index = task, batch, ready, slot, index


class PretrainStepper:
    def __init__(self, student, teacher=None, kdl=None, lr=5e-5, betas=(0.9, 0.98), weight_decay=0.01,
                 max_grad_norm=5.0, use_graphs=False, rw_generator=None, side_stream=True,
                 branch_streams=True, co_update=False, t_lr=None, max_graphs=16, overlap=True,
                 pipeline_teacher=True, teacher_sm_budget=0, pdl=None, tasks=None):
        """co_update=True is ICoD (`--train_kdl_teacher`, agent_base.py:260-279): the teacher is trained too, from
        the s2t losses, with its own arena / AdamW state / clip, and both models step once per batch."""
        # a model the caller already wrapped like the reference does (wrap_model, utils/misc.py:57-71) is unwrapped: the
        # stepper exchanges the flat gradient arena itself and never calls forward through the DDP reducer
        DDP = torch.nn.parallel.DistributedDataParallel
        student = student.module if isinstance(student, DDP) else student
        teacher = teacher.module if isinstance(teacher, DDP) else teacher
        self.student, self.teacher = student, teacher
        self.kdl = makd.kdl_config(kdl)
        self.co_update = bool(co_update and teacher is not None)
        lowp = student.compute_dtype == torch.bfloat16
        self.arena = ParamArena(student, lowp=lowp)
        self.t_arena = self.t_opt = None
        if self.co_update:
            self.t_arena = ParamArena(teacher, lowp=teacher.compute_dtype == torch.bfloat16)
            self.t_opt = FusedAdamW(self.t_arena, lr=lr if t_lr is None else t_lr, betas=betas,
                                    weight_decay=weight_decay, max_grad_norm=max_grad_norm)
        elif teacher is not None:
            if lowp:
                self.t_arena = ParamArena(teacher, lowp=True, requires_grad_only=False, with_grads=False)
            for p in teacher.parameters():
                p.requires_grad_(False)
        self.opt = FusedAdamW(self.arena, lr=lr, betas=betas, weight_decay=weight_decay, max_grad_norm=max_grad_norm)
        # task activity: with `tasks` (the tasks the loop alternates between) the optimizer leaves the parameters a
        # step's task never touches alone -- heads of the other tasks, unused KD projections -- and counts their steps
        # separately, as the reference AdamW does for parameters whose grad is None (optim/adamw.py:66-67, :86)
        self.tasks = tuple(tasks) if tasks else None
        if self.tasks:
            from .model import inactive_in_task
            kd = teacher is not None
            self.opt.configure_tasks(self.tasks, lambda t, n: inactive_in_task(t, n, kd=kd))
            if self.t_opt is not None:
                self.t_opt.configure_tasks(self.tasks, lambda t, n: inactive_in_task(t, n, kd=False))
        if teacher is not None:
            # MAKD reads the attention maps of both models on their common depth only (agent.py:560,654,671; makd.py):
            # the deeper model does not produce the maps nobody compares
            cs, ct = student.config, teacher.config
            n_txt = min(cs.num_l_layers, ct.num_l_layers)
            n_x = min(cs.num_x_layers, ct.num_x_layers, n_txt)
            student.set_kd_attn_depth(n_txt, n_x)
            teacher.set_kd_attn_depth(n_txt, n_x)
        self.use_graphs = use_graphs
        ops.enable_side_stream(side_stream)
        ops.enable_branch_streams(branch_streams)
        self.graphs = {}
        self.max_graphs = int(max_graphs)  # shape signatures beyond this many run eagerly (no unbounded cache)
        self._graph_overflow_warned = False
        self.rw_generator = rw_generator
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.device = self.arena.device
        self.allreduce = FlatAllReduce(self.arena.flat_g) if self.world > 1 else None
        self.t_allreduce = FlatAllReduce(self.t_arena.flat_g) if (self.world > 1 and self.co_update) else None
        if self.world > 1:
            broadcast_flat(self.arena.flat_p, 0)  # DDP's wrap-time parameter broadcast (utils/misc.py:63-66)
            self.arena.refresh_lowp()
            if self.co_update:
                broadcast_flat(self.t_arena.flat_p, 0)
                self.t_arena.refresh_lowp()
        # gradient exchange overlapped with backward (DDP's bucketed overlap, utils/misc.py:57-71): the model marks the
        # points of backward at which a group of layers is complete (model._mark), parallel.StageSync starts that
        # group's all-reduce there; the rest follows when backward has been issued
        # (below ~64 MB of gradients the extra collectives and event waits cost more than they hide: measured at
        # 2 GPUs, 39 MB arena: 2.17 ms/step plain vs 2.25 overlapped; 1.4 GB arena: 19.7 vs 19.0)
        big = self.arena.total * 4 + (self.t_arena.total * 4 if self.co_update else 0) >= (64 << 20)
        self.overlap = (overlap == "force" or (bool(overlap) and big)) and self.world > 1
        self.syncs = []
        self.comm_stream = None
        if self.overlap:
            self.comm_stream = torch.cuda.Stream()
            pairs = [(student, self.arena)] + ([(teacher, self.t_arena)] if self.co_update else [])
            for m, a in pairs:
                sy = StageSync(a, len(m.bert.lang_encoder.layer))
                m.bert.stage_cb = sy
                self.syncs.append(sy)
        # frozen teacher + CUDA graphs: the teacher's forward is its own graph on its own stream, and the forward of
        # the NEXT batch (announced through step(..., next=)) runs while this batch's student back-propagates.  The
        # teacher does not depend on the student's update, so every result is identical to the serial order.
        self.pipeline_teacher = bool(pipeline_teacher) and use_graphs and teacher is not None and not self.co_update
        self._t_stream = None
        self._t_inflight = None
        # SMs the teacher graph's persistent GEMMs may occupy (0 = all): the teacher graph runs beside the student
        # graph, and a persistent CTA holds an SM's whole shared memory and TMEM for the length of its kernel
        self.teacher_sm_budget = int(teacher_sm_budget)
        # programmatic dependent launch: a win for a lone chain of small kernels, a loss when graph branches of two
        # models compete for the SMs (an early-launched grid holds SM slots while it waits).  Default: on for the graph
        # that trains, off for the frozen teacher's own graph.  (student_pdl, teacher_pdl) or one bool for both.
        # measured (B200, distillation step, ms): student/teacher PDL on/off 5.08, off/off 5.15, on/on 5.35, off/on 5.52
        if pdl is None:
            pdl = (True, False)
        self.pdl = (bool(pdl), bool(pdl)) if not isinstance(pdl, (tuple, list)) else (bool(pdl[0]), bool(pdl[1]))
        self.launches_per_step = None
        self._copy_stream, self._staging = None, {}
        self._rw_dev = None
        self._next = None
        self._last_res = None

    KD_TASKS = ("mlm", "sap")

    def kd_task(self, task):
        """MAKD is defined for MLM and SAP steps (the `predict` ability needs the task's logits, the
        MKTD weights the teacher's per-sample task loss; model.inactive_in_task encodes the same rule for the
        optimizer).  The other tasks of a run that has a teacher -- `cfp` in the reference's own task list
        (config/r2r_magic_pretrain.json:49-53), `mrc`, `og` -- take the plain supervised step."""
        return self.teacher is not None and task[:3] in self.KD_TASKS

    @staticmethod
    def _detach_res(res):
        """Values only: holding the loss tensors themselves would keep the step's autograd graph (and its
        AccumulateGrad nodes, bound to the stream they were created on) alive into the next capture."""
        if res is None:
            return None
        return dict(per_seg=res["per_seg"].detach() if res["per_seg"] is not None else None, owner=list(res["owner"]),
                    kl=res["kl"].detach() if res["kl"] is not None else None, mse_total=None)

    def last_named_losses(self):
        """The reference's 10 named MAKD scalars (agent.py:824-835) of the most recent step, for logging: one host
        sync.  Under CUDA graphs these are the captured graph's static output tensors, refreshed by every replay.
        None without a teacher."""
        res = getattr(self, "_last_res", None)
        return makd.named_losses(res) if res is not None else None

    def exchange_description(self):
        if self.world == 1:
            return None
        if self.overlap:
            mb = sum((hi - lo) * 4 for sy in self.syncs for st in sy.stages for lo, hi in st) / 1e6
            tot = sum(sy.arena.total * 4 for sy in self.syncs) / 1e6
            return (f"NCCL all-reduce(AVG) of the flat fp32 gradient arena in 64 MB buckets; {mb:.0f} of {tot:.0f} MB "
                    "(cross-modal encoders + heads, text layer groups from the top, panorama encoder) start DURING backward from in-graph event "
                    "nodes, the rest when backward has been issued")
        n = len(self.allreduce.bounds) + (len(self.t_allreduce.bounds) if self.t_allreduce is not None else 0)
        return f"NCCL all-reduce(AVG) of the flat fp32 gradient arena, {n} buckets, issued after backward"

    # -- the device side of one step -------------------------------------------------------------
    def _finish(self, fired=None, task=None):
        """Gradient exchange + optimizer (kept outside the captured graph when world > 1)."""
        if self.overlap:
            # what backward did not already exchange; the teacher's exchange (ICoD) runs under the student's optimizer
            for i, sy in enumerate(self.syncs):
                sy.issue_rest(sy.fired if fired is None else fired[i])
            self.syncs[0].wait()
            self.opt.apply(task)
            if self.co_update:
                self.syncs[1].wait()
                self.t_opt.apply(task)
            return
        # both exchanges are issued up front: the teacher's all-reduce (ICoD) runs under the student's optimizer
        if self.allreduce is not None:
            self.allreduce.start()
        if self.co_update and self.t_allreduce is not None:
            self.t_allreduce.start()
        if self.allreduce is not None:
            self.allreduce.finish()
        self.opt.apply(task)
        if self.co_update:  # agent_base.py:271-274: clip + step the student, then clip + step the teacher
            if self.t_allreduce is not None:
                self.t_allreduce.finish()
            self.t_opt.apply(task)

    def _device_step(self, task, batch, rw, finish=True, mode="eager", t_out=None):
        for sy in self.syncs:
            sy.begin(mode)
        self.arena.zero_grad()
        kd = self.kd_task(task)
        if self.co_update:
            self.t_arena.zero_grad()
            if kd:
                # agent.py:869-871: under RW the teacher's ability weights ARE the student's draw of this step
                mix, mix_t, res, _, _, _ = makd.icod_step_loss(self.student, self.teacher, batch, task, rw, rw,
                                                               self.kdl)
            else:  # a task MAKD is not defined for: each model takes its own supervised step
                res = None
                mix, mix_t = (ops.loss_mix(None, None, o["loss"], 0.0, o.get("loss_inv_n"))
                              for o in (self.student(batch, task, True), self.teacher(batch, task, True)))
            self._last_res = self._detach_res(res)
            # the reference calls loss.backward(retain_graph=True) then t_loss.backward() (agent_base.py:260-268);
            # every cross-model target is detached, so the two graphs are disjoint and one pass over their sum
            # produces exactly those gradients
            (mix[0] + mix_t[0]).backward()
            ops.join_side_stream()
            for sy in self.syncs:
                sy.join_marker()
            if finish:
                self._finish(task=task)
            return torch.cat([mix, mix_t]).detach()
        if kd and t_out is not None:  # teacher outputs computed by the teacher's own graph
            mix, res, s_out = makd.student_distill_loss(self.student, t_out, batch, task, rw, self.kdl)
        elif kd:
            mix, res, s_out, t_out = makd.distill_step_loss(self.student, self.teacher, batch, task, rw, self.kdl)
        else:  # no teacher, or a task MAKD is not defined for (kd_task): the supervised step
            res = None
            s_out = self.student(batch, task, True)
            mix = ops.loss_mix(None, None, s_out["loss"], 0.0, s_out.get("loss_inv_n"))
        self._last_res = self._detach_res(res)
        mix[0].backward()
        ops.join_side_stream()
        for sy in self.syncs:
            sy.join_marker()
        if finish:
            self._finish(task=task)
        # values only: a tensor with a grad_fn would keep this step's autograd graph alive, and with it the per-
        # parameter AccumulateGrad nodes bound to the streams of THIS step -- a later capture whose ops run on other
        # streams would then be invalidated by the engine's sync with those (uncaptured) streams
        return mix.detach()

    # -- host -> device prefetch (the reference's PrefetchLoader, data/loader.py:78-124) -----------------
    def prefetch(self, task, host_batch):
        """Start copying a pinned host batch (graph_index.prepare_batch output) to the device on a dedicated
        copy stream and return a handle that `step()` accepts in place of a device batch.  Two staging sets per
        batch signature are used alternately, so the copy of batch i+1 overlaps the compute of batch i."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
        sig = self._signature(task, host_batch)
        slot = self._staging.get(sig)
        if slot is None:
            slot = self._staging[sig] = dict(sets=[alloc_like(host_batch, self.device) for _ in range(2)],
                                             free=[None, None], n=0)
            # the fresh buffers may recycle memory that kernels already queued on this stream still use
            self._copy_stream.wait_stream(torch.cuda.current_stream())
        i = slot["n"] % 2
        slot["n"] += 1
        dst = slot["sets"][i]
        with torch.cuda.stream(self._copy_stream):
            if slot["free"][i] is not None:
                self._copy_stream.wait_event(slot["free"][i])  # the previous consumer of this set has finished
            copy_batch_(dst, host_batch)  # ONE copy when the host batch is flat (graph_index.flatten_batch)
            for k, v in host_batch.items():  # python-side members (vp-id lists, static ints) travel by reference
                if not torch.is_tensor(v) and k != INDEX_KEY:
                    dst[k] = v
            ready = torch.cuda.Event()
            ready.record(self._copy_stream)
        return _Prefetched(task, dst, ready, slot, i)

    def step(self, task, batch, lr=None, next=None):
        """batch: device batch with index tables (graph_index.prepare_batch + batch_to_device), or the handle
        returned by `prefetch()`.  Returns the device tensor [total, supervised_mean, kd_total] (no host sync).
        `next` = (task, batch or prefetch handle) of the FOLLOWING step, if known (a prefetching loader knows it):
        with a frozen teacher and CUDA graphs its teacher forward is started now, under this step's backward.  A
        device batch announced this way must already be resident (not still being written on the current stream)."""
        handle = None
        if isinstance(batch, _Prefetched):
            handle, batch = batch, batch.batch
            torch.cuda.current_stream().wait_event(handle.ready)
        nxt = None
        if next is not None and self.pipeline_teacher:
            t2, b2 = next
            nxt = (t2, b2.batch, b2.ready) if isinstance(b2, _Prefetched) else (t2, b2, None)
        self._next = nxt
        out = self._step(task, batch, lr)
        if handle is not None:
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream())
            handle.slot["free"][handle.index] = done
        return out

    def _step(self, task, batch, lr=None):
        rw = None
        if self.teacher is not None:
            if self.use_graphs:  # device-resident weights: the captured kernels read this step's draw
                if self._rw_dev is None:
                    self._rw_dev = torch.ones(5, dtype=torch.float32, device=self.device)
                    self._rw_ring = ops.PinnedRing(5)
                rw = makd.mkrw_weights(self.kdl["rw_temp"], generator=self.rw_generator, out=self._rw_dev,
                                       ring=self._rw_ring)
            else:
                rw = makd.mkrw_weights(self.kdl["rw_temp"], generator=self.rw_generator)
        self.opt.set_hyper(lr, task)
        if self.co_update:
            self.t_opt.set_hyper(None, task)
        ops.bump_seed(self.device)
        for a in (self.arena, self.t_arena):  # parameters written through torch since the last step (resume, EMA)
            if a is not None:
                a.sync_lowp()
        if self.use_graphs:
            sig = self._signature(task, batch)
            if sig in self.graphs or len(self.graphs) < self.max_graphs:
                return self._graph_step(task, batch, rw, sig)
            if not self._graph_overflow_warned:
                self._graph_overflow_warned = True
                import warnings
                warnings.warn(f"PretrainStepper: more than {self.max_graphs} batch shape signatures; further shapes run "
                              "eagerly -- pad batches to fixed capacities (graph_index.pad_batch) to replay one graph "
                              "per task")
        return self._device_step(task, batch, rw)

    # -- CUDA-graph replay ------------------------------------------------------------------------
    @staticmethod
    def _signature(task, batch):
        sig = [task]
        for k in sorted(batch.keys()):
            v = batch[k]
            if k == FLAT_KEY:
                continue
            if torch.is_tensor(v):
                sig.append((k, tuple(v.shape)))
            elif k == INDEX_KEY:
                sig.append(tuple((kk, tuple(vv.shape) if torch.is_tensor(vv) else vv) for kk, vv in sorted(v.items())))
        return tuple(sig)

    def _capture_teacher(self, task, batch):
        """The frozen teacher's forward as its own graph: static inputs, static outputs (the dict the student graph
        reads), two events -- `ready` (outputs valid) and `done` (the student step that read them has been issued)."""
        static = alloc_like(batch, self.device)
        copy_batch_(static, batch)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        lib = _lib.load()
        lib.magic_gemm_set_sm_budget(self.teacher_sm_budget)
        lib.magic_set_pdl(int(self.pdl[1]))
        try:
            with torch.cuda.stream(s):
                makd.teacher_forward(self.teacher, static, task)  # warm-up (allocator, lazy init)
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            n0 = _lib.COUNTERS["launches"]
            import os as _os
            tp = _os.environ.get("MAGIC_TEACHER_PRIO", "0") == "1"
            with (torch.cuda.graph(g, stream=torch.cuda.Stream(priority=-1)) if tp else torch.cuda.graph(g)):
                t_out = makd.teacher_forward(self.teacher, static, task)
        finally:
            lib.magic_gemm_set_sm_budget(0)
            lib.magic_set_pdl(-1)
        return dict(g=g, static=static, t_out=t_out, n=_lib.COUNTERS["launches"] - n0, ready=torch.cuda.Event(),
                    done=torch.cuda.Event())

    def _launch_teacher(self, t_ent, batch, ready=None, serial=False):
        if self._t_stream is None:
            self._t_stream = torch.cuda.Stream()
        T = self._t_stream
        if serial:  # not announced in advance: the batch may still be in flight on the current stream
            T.wait_stream(torch.cuda.current_stream())
        if ready is not None:
            T.wait_event(ready)          # host -> device copy of a prefetched batch
        T.wait_event(t_ent["done"])      # the previous step that read these outputs (a never-recorded event is a no-op)
        with torch.cuda.stream(T):
            copy_batch_(t_ent["static"], batch)
            t_ent["g"].replay()
            t_ent["ready"].record(T)
        _lib.COUNTERS["launches"] += t_ent["n"]

    def _graph_step(self, task, batch, rw, sig=None):
        sig = self._signature(task, batch) if sig is None else sig
        entry = self.graphs.get(sig)
        if entry is None:
            dev = self.device
            t_ent = None
            self._t_inflight = None
            if self.pipeline_teacher and self.kd_task(task):
                t_ent = self._capture_teacher(task, batch)
                t_ent["g"].replay()  # valid teacher outputs for the student's warm-up runs and capture
            t_out = t_ent["t_out"] if t_ent is not None else None
            static = alloc_like(batch, dev)
            copy_batch_(static, batch)
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            # warm-up on a side stream (allocator, cudaFuncSetAttribute, lazy init) WITHOUT changing the
            # training state: parameters and moments are restored afterwards
            pairs = [(self.arena, self.opt)] + ([(self.t_arena, self.t_opt)] if self.co_update else [])
            snap = [(a.flat_p.clone(), o.m.clone(), o.v.clone()) for a, o in pairs]
            with torch.cuda.stream(s):
                for _ in range(2):
                    self._device_step(task, static, rw, finish=self.world == 1, mode=None, t_out=t_out)
                    if self.world > 1:
                        for _, o in pairs:
                            o.apply(task)
                for (a, o), (p0, m0, v0) in zip(pairs, snap):
                    a.flat_p.copy_(p0)
                    o.m.copy_(m0)
                    o.v.copy_(v0)
                    a.refresh_lowp()
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            n0 = _lib.COUNTERS["launches"]
            _lib.load().magic_set_pdl(int(self.pdl[0]))
            # kernel nodes inherit the priority of the stream they are captured on: the training graph (a latency-bound
            # chain of small kernels) is captured at high priority, the frozen teacher's throughput-bound graph at the
            # default one, so a student CTA never queues behind the teacher's pending CTAs when the two graphs run
            # side by side (MAGIC_STUDENT_PRIO=0 switches this off)
            import os as _os
            hp = self.pipeline_teacher and _os.environ.get("MAGIC_STUDENT_PRIO", "0") == "1"
            cap_stream = torch.cuda.Stream(priority=-1) if hp else None
            try:
                with (torch.cuda.graph(g, stream=cap_stream) if cap_stream is not None else torch.cuda.graph(g)):
                    out = self._device_step(task, static, rw, finish=self.world == 1, mode="capture", t_out=t_out)
            finally:
                _lib.load().magic_set_pdl(-1)
            marks = [(list(sy.fired), dict(sy.events)) for sy in self.syncs]
            entry = (g, static, out, _lib.COUNTERS["launches"] - n0, marks, t_ent, self._last_res)
            self.graphs[sig] = entry
        g, static, out, n_launch, marks, t_ent, self._last_res = entry
        cur = torch.cuda.current_stream()
        if t_ent is not None:
            if self._t_inflight != (sig, id(batch)):
                self._launch_teacher(t_ent, batch, serial=True)
            self._t_inflight = None
            cur.wait_event(t_ent["ready"])
        copy_batch_(static, batch)
        g.replay()
        _lib.COUNTERS["launches"] += n_launch  # kernels replayed inside the graph
        if t_ent is not None:
            t_ent["done"].record(cur)
        if self._next is not None and self.pipeline_teacher:
            # the next batch's teacher forward starts now, under this step's backward (also when this step itself had
            # no teacher: a cfp step between two distillation steps)
            t2, b2, ready2 = self._next
            sig2 = self._signature(t2, b2)
            e2 = self.graphs.get(sig2)
            if e2 is not None and e2[5] is not None:
                self._launch_teacher(e2[5], b2, ready=ready2)
                self._t_inflight = (sig2, id(b2))
        if self.world > 1:
            # the exchange of every stage that completed inside the graph starts at its in-graph event
            for sy, (fired, events) in zip(self.syncs, marks):
                sy.after_replay(self.comm_stream, fired, events)
            self._finish([f for f, _ in marks] if self.overlap else None, task=task)
        return out
