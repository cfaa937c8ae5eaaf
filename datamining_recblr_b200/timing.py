"""Device-side timing helpers shared by bench.py and tools/ (CUDA events on the current stream)."""
import statistics

import torch

_FLUSH = {}


def flush_l2(device=None):
    """Evicts L2 by writing a buffer larger than the 126 MB L2 (outside any timed region)."""
    device = device or torch.cuda.current_device()
    buf = _FLUSH.get(device)
    if buf is None:
        buf = torch.empty(256 << 20, dtype=torch.uint8, device=device)
        _FLUSH[device] = buf
    buf.zero_()


def time_cuda(fn, warmup=3, iters=10, flush=True):
    """Runs fn() warmup+iters times; returns per-iteration device times in ms (CUDA events around each call,
    L2 flushed between calls when flush=True)."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        if flush:
            flush_l2()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        times.append(s.elapsed_time(e))
    return times


def summarize(times):
    return {"median_ms": statistics.median(times), "min_ms": min(times), "n": len(times)}
