"""ctypes binding of libbdlru.so (include/bdlru.h).  There is no fallback: if the library is missing the
import fails loudly, and every op raises on non-CUDA tensors."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# BDLRU_LIB selects another build of the SAME library (tuning variants from tools/ce_variants.py); never a fallback
LIB_PATH = os.environ.get("BDLRU_LIB") or os.path.join(_HERE, "libbdlru.so")

F32, BF16 = 0, 1
ABI_VERSION = 4


class View(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("bstride", ctypes.c_int64), ("rstride", ctypes.c_int64)]


class BdlruError(RuntimeError):
    pass


_p, _i, _i64, _sz, _f, _u64 = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t, ctypes.c_float,
                               ctypes.c_uint64)

# name -> (restype, argtypes); mirrors include/bdlru.h declaration by declaration
SIGNATURES = {
    "bdlru_version": (_i, []),
    "bdlru_last_error": (ctypes.c_char_p, []),
    "bdlru_launch_count": (_u64, []),
    "bdlru_build_info": (ctypes.c_char_p, []),
    "bdlru_scan_fwd": (_i, [_p, _p, _p, _i64, _i64, _p]),
    "bdlru_scan_bwd": (_i, [_p, _p, _p, _p, _p, _i64, _i64, _p]),
    "bdlru_gated_scan_fwd": (_i, [View, View, View, _p, _p, _i64, View, View, View, _i, _i, _i, _i, _p]),
    "bdlru_gated_scan_bwd_workspace_bytes": (_sz, [_i, _i, _i]),
    "bdlru_gated_scan_bwd": (_i, [View, View, View, _p, _p, _i64, View, View, View, View, View, View, View,
                                  _p, _p, _p, _sz, _i, _i, _i, _i, _p]),
    "bdlru_phantom_h0_fwd": (_i, [_p, _p, _p, _p, _i, _i, _p, _p, _p]),
    "bdlru_phantom_h0_bwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p, _p]),
    "bdlru_scan_cl_fwd": (_i, [View, View, _p, _i64, View, _i, _i, _i, _i, _p]),
    "bdlru_scan_cl_bwd": (_i, [View, _p, _i64, View, View, View, View, _p, _p, _sz, _i, _i, _i, _i, _p]),
    "bdlru_conv1d_fwd": (_i, [View, _p, _p, View, _i, _i, _i, _i, _i, _i, _p]),
    "bdlru_conv1d_bwd_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "bdlru_conv1d_bwd": (_i, [View, _p, _p, View, View, _p, _p, _p, _sz, _i, _i, _i, _i, _i, _i, _p]),
    "bdlru_embed_ln_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i, _f, _f, _u64, _p, _i, _i, _p]),
    "bdlru_embed_ln_bwd_workspace_bytes": (_sz, [_i64, _i]),
    "bdlru_embed_ln_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _i64, _i64, _i, _f, _u64, _p, _i64, _i, _i, _p]),
    "bdlru_embed_ln_bwd_rows": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _i64, _i64, _i, _f, _u64, _p, _i64, _i, _i, _p]),
    "bdlru_scatter_add_rows": (_i, [_p, _p, _i64, _i, _i, _i64, _i64, _i64, _p, _p]),
    "bdlru_add_ln_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _i64, _i, _f, _f, _u64, _p, _i, _p]),
    "bdlru_add_ln_bwd_workspace_bytes": (_sz, [_i64, _i]),
    "bdlru_add_ln_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _i64, _i, _f, _u64, _p, _i, _p]),
    "bdlru_silu_dropout_fwd": (_i, [_p, _p, _i64, _f, _u64, _p, _i, _p]),
    "bdlru_silu_dropout_bwd": (_i, [_p, _p, _p, _i64, _f, _u64, _p, _i, _p]),
    "bdlru_colsum_workspace_bytes": (_sz, [_i64, _i]),
    "bdlru_colsum": (_i, [_p, _i64, _i, _i64, _i, _p, _p, _sz, _p]),
    "bdlru_core_fwd_supported": (_i, [_i, _i]),
    "bdlru_core_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "bdlru_layer_fwd_supported": (_i, [_i, _i, _i, _i]),
    "bdlru_layer_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _f, _p, _i, _i, _i, _i, _p]),
    "bdlru_table_adam_step": (_i, [_p, _p, _p, _p, _p, _i64, _f, _f, _f, _f, _f, _i64, _p]),
    "bdlru_fullsort_available": (_i, []),
    "bdlru_fullsort_topk_workspace_bytes": (_sz, [_i64, _i64, _i, _i]),
    "bdlru_fullsort_topk": (_i, [_p, _p, _i64, _i64, _i, _i, _i64, _i64, _p, _p, _p, _sz, _p]),
    "bdlru_topk_merge": (_i, [_p, _p, _i64, _i, _i, _p, _p, _p]),
    "bdlru_topk_merge_strided": (_i, [_p, _p, _i64, _i, _i, _i64, _i64, _p, _p, _p]),
    "bdlru_fullsort_ce_workspace_bytes": (_sz, [_i64, _i64, _i]),
    "bdlru_fullsort_ce_fwd": (_i, [_p, _p, _p, _i64, _i64, _i, _i64, _p, _p, _p, _p, _sz, _p]),
    "bdlru_fullsort_rowmax_workspace_bytes": (_sz, [_i64, _i64, _i]),
    "bdlru_fullsort_rowmax": (_i, [_p, _p, _i64, _i64, _i, _i, _p, _p, _sz, _p]),
    "bdlru_fullsort_ce_fwd_dq_workspace_bytes": (_sz, [_i64, _i64, _i]),
    "bdlru_fullsort_ce_fwd_dq": (_i, [_p, _p, _p, _i64, _i64, _i, _p, _p, _p, _sz, _p]),
    "bdlru_fullsort_ce_bwd": (_i, [_p, _p, _p, _p, _f, _p, _i64, _i64, _i, _i64, _p, _p, _p, _sz, _p]),
}

_lib = None
MISSING = []
# name of a C-ABI entry point -> list of (start, end) CUDA events; filled only while bench.py asks for it
# (kernel_timer below), so the per-kernel launch durations behind `roofline.achieved` are measured live, on the
# stream the kernels are launched on, inside the timed region.
TIMERS = {}


class _Timed:
    """Callable proxy for one entry point: records a CUDA-event pair around the call when its name is in TIMERS."""
    __slots__ = ("name", "fn")

    def __init__(self, name, fn):
        self.name, self.fn = name, fn

    def __call__(self, *args):
        rec = TIMERS.get(self.name)
        if rec is None:
            return self.fn(*args)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = self.fn(*args)
        e.record()
        rec.append((s, e))
        return rc


class _Lib:
    pass


def load():
    """Loads libbdlru.so once; raises if it has not been built (python -m datamining_recblr_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BdlruError(f"{LIB_PATH} not found: build it with `python -m datamining_recblr_b200.build` "
                         "(there is no CPU or PyTorch fallback for this path)")
    cdll = ctypes.CDLL(LIB_PATH)
    lib = _Lib()
    lib.cdll = cdll
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(cdll, name)
        except AttributeError:
            MISSING.append(name)  # tests/test_abi.py requires this list to be empty
            continue
        fn.restype, fn.argtypes = res, args
        setattr(lib, name, _Timed(name, fn) if res is _i and args else fn)
    if lib.bdlru_version() != ABI_VERSION:
        raise BdlruError(f"libbdlru.so ABI {lib.bdlru_version()} != expected {ABI_VERSION}: rebuild")
    _lib = lib
    return lib


def kernel_timer(names):
    """Start recording per-call CUDA-event pairs for the given entry points; returns the dict to read back."""
    TIMERS.clear()
    for n in names:
        TIMERS[n] = []
    return TIMERS


def kernel_timer_stop():
    """Stops recording; returns {name: [ms, ...]} (synchronises)."""
    torch.cuda.synchronize()
    out = {n: [s.elapsed_time(e) for s, e in ev] for n, ev in TIMERS.items()}
    TIMERS.clear()
    return out


def build_info():
    """{'abi': int, 'digest': str, 'tuning': bool, 'matches_sources': bool} of the loaded library."""
    from . import build as B
    info = load().bdlru_build_info().decode()
    dig = info.split("src=")[1].split()[0]
    return {"abi": int(load().bdlru_version()), "digest": dig[:16], "tuning": info.endswith(" tuning"),
            "matches_sources": dig == B.source_digest()}


def check(rc):
    if rc != 0:
        raise BdlruError(f"libbdlru error {rc}: {load().bdlru_last_error().decode()}")


def launch_count():
    return int(load().bdlru_launch_count())


def dtype_tag(t: torch.Tensor):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise BdlruError(f"unsupported dtype {t.dtype} (float32 or bfloat16)")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise BdlruError("this op runs on CUDA tensors only (no CPU fallback)")


def stream_ptr(t: torch.Tensor):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


NULL_VIEW = View(None, 0, 0)


def view3(t):
    """bdlru_view of a [B, T, C] tensor whose channel stride is 1 (any batch/row strides)."""
    if t is None:
        return NULL_VIEW
    assert t.dim() == 3 and (t.stride(2) == 1 or t.shape[2] == 1), "channel-last view required"
    return View(t.data_ptr(), t.stride(0), t.stride(1))


def cl_ok(t):
    """True if t [B,T,C] can be passed as a view without a copy (unit channel stride, 4-element alignment)."""
    es = t.element_size()
    return (t.stride(2) == 1 and t.stride(0) % 4 == 0 and t.stride(1) % 4 == 0 and t.data_ptr() % (4 * es) == 0)


def as_cl(t):
    return t if cl_ok(t) else t.contiguous()
