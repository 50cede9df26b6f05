"""ctypes binding of libmagic_b200.so (the C ABI declared in include/magic_b200.h).

The argtypes of every entry point are derived from the header itself, so the header is the single
source of truth for the boundary.  There is NO fallback: if the shared library is missing the import
of any compute op raises (the product path must fail loudly without its CUDA extension).
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "magic_b200.h")
LIB_PATH = os.path.join(_HERE, "lib", "libmagic_b200.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
MAKD_MAX_SEGS = 32

_CT = {
    "int": ctypes.c_int, "long": ctypes.c_long, "long long": ctypes.c_longlong, "float": ctypes.c_float,
    "unsigned": ctypes.c_uint, "cudaStream_t": ctypes.c_void_p,
}


class MagicMseSeg(ctypes.Structure):
    _fields_ = [("s", ctypes.c_void_p), ("t", ctypes.c_void_p), ("ds", ctypes.c_void_p), ("w", ctypes.c_void_p),
                ("scale_dev", ctypes.c_void_p), ("rows", ctypes.c_longlong), ("inner", ctypes.c_longlong), ("s_rs", ctypes.c_longlong),
                ("t_rs", ctypes.c_longlong), ("scale", ctypes.c_float), ("s_dt", ctypes.c_int),
                ("t_dt", ctypes.c_int), ("vec_ok", ctypes.c_int)]


def parse_header(path=HEADER):
    """-> {name: (restype, [argtypes])} for every function the header declares."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"^(const char\*|int)\s+(magic_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.M | re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        argtypes = []
        for a in [x.strip() for x in args.replace("\n", " ").split(",")]:
            if a in ("void", ""):
                continue
            if "*" in a:
                argtypes.append(ctypes.c_void_p)
                continue
            toks = a.split()
            ty = " ".join(toks[:-1]) if len(toks) > 1 else toks[0]
            ty = ty.replace("const ", "").strip()
            argtypes.append(_CT[ty])
        decls[name] = (ctypes.c_char_p if ret.startswith("const char") else ctypes.c_int, argtypes)
    return decls


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"libmagic_b200.so not found at {LIB_PATH}: build it with `python -c 'import __graft_entry__ as g; "
            f"g.build()'` (or python vln-magic_b200/build.py). There is no CPU / PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (ret, argtypes) in parse_header().items():
        fn = getattr(lib, name)  # AttributeError if the header declares something the library lacks
        fn.restype = ret
        fn.argtypes = argtypes
    _lib = lib
    return lib


class MagicError(RuntimeError):
    pass


# kernels launched per C-ABI call (for the launch counter bench.py reports as gpu_launches)
_LAUNCHES = {"magic_attn_bwd": 2, "magic_scatter_rows": 1, "magic_gmap_aggregate_bwd": 1}
ERR_UNSUPPORTED = 3
COUNTERS = {"calls": 0, "launches": 0}
_PROFILE = None  # {name: [record]} when bench.py profiles kernel families
_PROFILE_EVERY = 0  # > 0: keep the stream busy with a delay kernel every N calls so the host stays ahead of the GPU
_PROFILE_N = 0
_PROFILE_GRAPH = False  # True: brackets are external event-record NODES of the graph being captured
_PROFILE_TAG = None


class GraphEvent:
    """A CUDA event of libmagic_b200 (magic_event_*): recorded with cudaEventRecordExternal while a stream is being
    captured, so it becomes a node of the graph and is re-stamped by every replay."""
    __slots__ = ("h",)

    def __init__(self):
        h = ctypes.c_void_p()
        rc = load().magic_event_create(ctypes.byref(h))
        if rc != 0:
            raise MagicError(f"magic_event_create failed: {load().magic_last_error().decode()}")
        self.h = h

    def record(self):
        rc = load().magic_event_record(self.h, stream())
        if rc != 0:
            raise MagicError(f"magic_event_record failed: {load().magic_last_error().decode()}")

    def elapsed_time(self, other):
        ms = ctypes.c_float()
        rc = load().magic_event_elapsed_ms(self.h, other.h, ctypes.byref(ms))
        if rc != 0:
            raise MagicError(f"magic_event_elapsed_ms failed: {load().magic_last_error().decode()}")
        return ms.value


def profile_start(delay_every=0, delay_cycles=3e6, graph=False):
    """Per-call CUDA-event timing of every C-ABI call (bench.py's roofline pass).

    graph=True (the default pass of bench.py): nothing is recorded in eager execution; while a stream is being
    CAPTURED every call is bracketed by two external event-record nodes, so the captured graph carries its own
    stopwatch and each replay yields the duration of every kernel of the replayed step -- same graph topology
    (stream branches) as the timed region; an event node between two kernels removes their programmatic-dependent-
    launch overlap, which is the only difference.

    graph=False: eager mode.  The host issues launches more slowly than the GPU retires small kernels; a short
    `magic_delay` kernel every `delay_every` calls (outside the event brackets) lets the host queue the next calls
    while the stream is busy, so each bracket measures back-to-back device execution and not host launch latency."""
    global _PROFILE, _PROFILE_EVERY, _PROFILE_N, _PROFILE_CYCLES, _PROFILE_GRAPH
    _PROFILE = {}
    _PROFILE_GRAPH = bool(graph)
    _PROFILE_EVERY, _PROFILE_N, _PROFILE_CYCLES = int(delay_every), 0, int(delay_cycles)


def profile_stop():
    global _PROFILE, _PROFILE_GRAPH
    p, _PROFILE = _PROFILE, None
    _PROFILE_GRAPH = False
    return p


def profile_tag(tag):
    global _PROFILE_TAG
    _PROFILE_TAG = tag


def _profiled(lib, name, args):
    global _PROFILE_N
    if _PROFILE_GRAPH:
        if not torch.cuda.is_current_stream_capturing():
            return getattr(lib, name)(*args)
        e0, e1 = GraphEvent(), GraphEvent()
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        _PROFILE.setdefault(name, []).append((e0, e1, args, _PROFILE_TAG))
        return rc
    if _PROFILE_EVERY > 0 and name != "magic_delay":
        if _PROFILE_N % _PROFILE_EVERY == 0:
            lib.magic_delay(_PROFILE_CYCLES, stream())
        _PROFILE_N += 1
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = getattr(lib, name)(*args)
    e1.record()
    _PROFILE.setdefault(name, []).append((e0, e1, args, _PROFILE_TAG))
    return rc


def call(name, *args):
    lib = load()
    COUNTERS["calls"] += 1
    COUNTERS["launches"] += _LAUNCHES.get(name, 1)
    if _PROFILE is not None:
        rc = _profiled(lib, name, args)
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise MagicError(f"{name} failed (rc={rc}): {lib.magic_last_error().decode()}")


def call_rc(name, *args):
    """Like `call`, for entry points that may decline a shape: returns 0, or ERR_UNSUPPORTED when NOTHING was launched
    (the caller then takes the general entry point); any other status raises."""
    lib = load()
    if _PROFILE is not None:
        rc = _profiled(lib, name, args)
        if rc == ERR_UNSUPPORTED and _PROFILE.get(name):
            _PROFILE[name].pop()  # nothing was launched inside that bracket
    else:
        rc = getattr(lib, name)(*args)
    if rc == ERR_UNSUPPORTED:
        return rc
    if rc != 0:
        raise MagicError(f"{name} failed (rc={rc}): {lib.magic_last_error().decode()}")
    COUNTERS["calls"] += 1
    COUNTERS["launches"] += _LAUNCHES.get(name, 1)
    return 0


def ptr(t):
    return None if t is None else t.data_ptr()


def dt(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise MagicError(f"unsupported dtype {t.dtype}")


def stream():
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise MagicError("magic_b200 ops need CUDA tensors (no CPU fallback exists)")
