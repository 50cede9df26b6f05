"""Data-parallel plumbing (one process per GPU; torch.distributed over NCCL/NVLink, gloo on CPU for tests).

The path shards by samples only: every rank runs the same step on its own batches and the
single exchange is the gradient all-reduce.  What this module replaces / restates from the reference:
  * DDP(find_unused_parameters=True) bucketed all-reduce      pretrain_src/utils/misc.py:57-71
      -> one flat fp32 gradient arena, all-reduced in a few large async buckets (no graph traversal, no
         per-parameter hooks; parameters a task does not touch simply contribute zeros)
  * the per-step 1-int task-id broadcast from rank 0           pretrain_src/data/loader.py:56-59
      -> a seeded task schedule every rank derives locally (zero collectives)
  * DistributedSampler sharding                                pretrain_src/data/loader.py:148-150
"""
import os

import torch
import torch.distributed as dist


def init_distributed(device=None):
    """env:// rendezvous (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT), like utils/distributed.py:73."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo")
    return rank, world, local


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class FlatAllReduce:
    """Averaging all-reduce of a flat gradient buffer in `n_buckets` contiguous async pieces, issued from the
    END of the buffer backwards (the order in which backward finishes the gradients)."""

    def __init__(self, flat, bucket_bytes=64 << 20):
        self.flat = flat
        n = flat.numel()
        per = max(1, bucket_bytes // flat.element_size())
        self.bounds = []
        hi = n
        while hi > 0:
            lo = max(0, hi - per)
            self.bounds.append((lo, hi))
            hi = lo
        self.world = world_size()

    def start(self):
        """Issue the bucket all-reduces (asynchronously, on the backend's own stream)."""
        self._handles, self._avg = [], True
        if self.world == 1:
            return
        # NCCL averages inside the collective (no extra pass over the arena); gloo (CPU tests) has no AVG
        self._avg = dist.get_backend() == "nccl"
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        self._handles = [dist.all_reduce(self.flat[lo:hi], op=op, async_op=True) for lo, hi in self.bounds]

    def finish(self):
        """Order the current stream after the exchange started by `start()`."""
        for h in self._handles:
            h.wait()
        if self._handles and not self._avg:
            self.flat.mul_(1.0 / self.world)
        self._handles = []

    def __call__(self):
        self.start()
        self.finish()


# ---------------------------------------------------------------------------------------------------
# gradient exchange overlapped with backward (DDP's bucketed overlap, pretrain_src/utils/misc.py:57-71)
# ---------------------------------------------------------------------------------------------------
STAGE0_PREFIXES = ("bert.local_encoder.encoder.", "bert.global_encoder.encoder.", "bert.global_encoder.sprel_linear",
                   "bert.txt_emb_w", "bert.kdl_img_w", "bert.kdl_avg_img_w", "bert.global_cross_w",
                   "bert.local_cross_w", "bert.vp_txt_w", "bert.gmap_txt_w", "mlm_head.", "global_sap_head.",
                   "local_sap_head.", "sap_fuse_linear.", "image_classifier.", "cfp_", "og_head.")


def text_stage_cuts(n_text_layers):
    """Text-layer indices at which the model marks a backward stage, top first: the gradient of the activation ENTERING
    layer c is complete when layers c.. have finished backward.  9 layers -> [6, 3]; 6 -> [4, 2]; 2 -> [1]."""
    if n_text_layers < 2:
        return []
    step = max(1, -(-n_text_layers // 3))
    cuts, c = [], n_text_layers - step
    while c > 0:
        cuts.append(c)
        c -= step
    return cuts


PANO_PREFIXES = ("bert.img_embeddings.pano_encoder.", "bert.img_embeddings.adaptive_pano_attn.")


def stage_ranges(arena, n_text_layers):
    """-> ([ranges of stage 0, 1, ...], rest ranges) as (lo, hi) element offsets of the gradient arena.

    Stage 0 = every weight-decay-group parameter of the cross-modal encoders, the KD projections and the task heads:
    complete when the gradients of the cross encoders' inputs (text output, gmap input, vp input) are complete.
    Stages 1.. = groups of text layers from the top (text_stage_cuts): complete when the gradient of the activation
    entering the group's first layer is complete.  Last stage = the panorama encoder: complete when the gradient of its
    input is.  (The position / step embedding weights of the two cross encoders produce gmap / vp INPUTS, so their
    gradients arrive after the stage-0 point: they stay in the rest, with the biases and LayerNorm parameters of the
    no-decay group, the embeddings, the lowest text layers and the image embeddings.)"""
    cuts = text_stage_cuts(n_text_layers)
    groups, hi_l = [], n_text_layers
    for c in cuts:
        groups.append(tuple(f"bert.lang_encoder.layer.{i}." for i in range(c, hi_l)))
        hi_l = c
    stages = [[] for _ in range(2 + len(groups))]
    for n, p, o, k in arena.entries:
        if o >= arena.n_decay:
            continue
        st = None
        if n.startswith(STAGE0_PREFIXES):
            st = 0
        elif n.startswith(PANO_PREFIXES):
            st = 1 + len(groups)
        else:
            for gi, pre in enumerate(groups):
                if n.startswith(pre):
                    st = 1 + gi
                    break
        if st is None:
            continue
        hi = o + (k + 7) // 8 * 8  # the arena pads every entry to 8 elements
        r = stages[st]
        if r and r[-1][1] == o:
            r[-1] = (r[-1][0], hi)
        else:
            r.append((o, hi))
    covered = sorted(x for r in stages for x in r)
    rest, pos = [], 0
    for lo, hi in covered:
        if lo > pos:
            rest.append((pos, lo))
        pos = hi
    if pos < arena.total:
        rest.append((pos, arena.total))
    return stages, rest


class StageSync:
    """Receives the model's backward stage marks (model.GlocalTextPathCMT._mark) for ONE model / gradient arena and
    starts the all-reduce of a stage's arena ranges as soon as the stage is complete.

    mode 'eager'   : the hook orders the current stream after every helper stream and issues the NCCL all-reduce.
    mode 'capture' : the step is being captured into a CUDA graph.  NCCL is NOT captured; instead the hook adds an
                     external event-record node (on a marker stream that depends on every capturing stream) to the
                     graph.  After each replay the host makes the communication stream wait for that event and
                     issues the all-reduce there: the exchange starts in the middle of the running graph.
    mode None      : marks are ignored (warm-up, single GPU)."""

    def __init__(self, arena, n_text_layers, bucket_bytes=64 << 20):
        self.arena = arena
        self.stages, self.rest = stage_ranges(arena, n_text_layers)
        self.bucket = max(1, bucket_bytes // 4)
        self.mode = None
        self.pending = [set() for _ in self.stages]
        self.expected = [False] * len(self.stages)
        self.fired = []          # stages completed during the current backward, in order
        self.events = {}         # capture mode: stage -> GraphEvent
        self.handles = []
        self.marker = None
        self.main = None

    # -- called by the model ------------------------------------------------------------------------------
    def expect(self, stage, key):
        if self.mode is None or stage >= len(self.stages):
            return
        self.pending[stage].add(key)
        self.expected[stage] = True

    def fire(self, stage, key):
        if self.mode is None or stage >= len(self.stages):
            return
        self.pending[stage].discard(key)
        if self.expected[stage] and not self.pending[stage] and stage not in self.fired:
            self.fired.append(stage)
            self._ready(stage)

    # -- stepper interface ----------------------------------------------------------------------------------
    def begin(self, mode):
        self.mode = mode
        # the step's (capture) origin stream; None on the CPU (gloo tests of the bucket plan)
        self.main = torch.cuda.current_stream() if (mode is not None and torch.cuda.is_available()) else None
        self.pending = [set() for _ in self.stages]
        self.expected = [False] * len(self.stages)
        self.fired = []
        if mode == "capture":
            self.events = {}

    def _streams(self):
        from . import ops
        out = list(ops._BRANCH.get("pool", {}).values()) + list(ops._ATTN_STREAMS.values())
        if ops._SIDE["stream"] is not None:
            out.append(ops._SIDE["stream"])
        if self.main is not None:
            out.append(self.main)
        return out

    def _ready(self, stage):
        cur = torch.cuda.current_stream()
        if self.mode == "eager":
            for s in self._streams():
                cur.wait_stream(s)
            self._issue(stage)
            return
        from ._lib import GraphEvent
        if self.marker is None:
            self.marker = torch.cuda.Stream()
        mk = self.marker
        mk.wait_stream(cur)
        for s in self._streams():
            with torch.cuda.stream(s):
                capturing = torch.cuda.is_current_stream_capturing()
            if capturing:
                mk.wait_stream(s)  # an edge from the stream's last node; adds no work to that stream
        ev = GraphEvent()
        with torch.cuda.stream(mk):
            ev.record()            # external event-record node: re-stamped by every replay
        self.events[stage] = ev

    def join_marker(self):
        """Capture mode: the marker stream must rejoin the origin stream before the capture ends."""
        if self.marker is not None and self.mode == "capture" and self.events:
            torch.cuda.current_stream().wait_stream(self.marker)

    def _issue(self, stage_or_ranges):
        ranges = self.stages[stage_or_ranges] if isinstance(stage_or_ranges, int) else stage_or_ranges
        g = self.arena.flat_g
        avg = dist.get_backend() == "nccl"
        op = dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM
        for lo, hi in ranges:
            pos = lo
            while pos < hi:
                end = min(hi, pos + self.bucket)
                self.handles.append((dist.all_reduce(g[pos:end], op=op, async_op=True), pos, end, avg))
                pos = end

    def after_replay(self, comm_stream, fired, events):
        """Graph mode, called right after g.replay(): the communication stream waits for each stage's in-graph event
        and issues that stage's all-reduce (NCCL orders its own stream after the stream it is called from)."""
        from ._lib import load
        lib = load()
        with torch.cuda.stream(comm_stream):
            for st in fired:
                lib.magic_stream_wait_event(comm_stream.cuda_stream, events[st].h)
                self._issue(st)

    def issue_rest(self, fired):
        """Everything that was not exchanged during backward (call when backward has been issued completely)."""
        ranges = list(self.rest)
        for st in range(len(self.stages)):
            if st not in fired:
                ranges += self.stages[st]
        self._issue(sorted(ranges))

    def wait(self):
        world = world_size()
        for h, lo, hi, avg in self.handles:
            h.wait()
            if not avg:
                self.arena.flat_g[lo:hi].mul_(1.0 / world)
        self.handles = []


def broadcast_flat(flat, src=0):
    """DDP's parameter broadcast at wrap time (utils/misc.py:63-66) on the flat parameter arena."""
    if world_size() > 1:
        dist.broadcast(flat, src)


def task_schedule(seed, n_steps, tasks, ratios):
    """The MetaLoader's multinomial task sampling (data/loader.py:50-75) as a pre-agreed seeded schedule:
    every rank computes the same list, so no per-step broadcast is needed."""
    g = torch.Generator().manual_seed(int(seed))
    p = torch.tensor([float(r) for r in ratios])
    idx = torch.multinomial(p / p.sum(), n_steps, replacement=True, generator=g)
    return [tasks[i] for i in idx.tolist()]


def shard_indices(n, rank, world, seed=0, shuffle=True, drop_last=False):
    """DistributedSampler semantics (pads by wrapping so every rank gets the same count)."""
    if shuffle:
        g = torch.Generator().manual_seed(int(seed))
        order = torch.randperm(n, generator=g).tolist()
    else:
        order = list(range(n))
    if drop_last:
        total = n // world * world
        order = order[:total]
    else:
        total = (n + world - 1) // world * world
        order = order + order[: total - len(order)]
    return order[rank:total:world]
