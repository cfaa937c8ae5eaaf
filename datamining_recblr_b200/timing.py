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


L2_BYTES = 126 << 20


def time_graph(make_set, run, set_bytes, iters=7):
    """HBM-resident timing of one op without flushes inside the timed region: the op is captured into a CUDA graph that
    runs it on R rotating input sets whose total footprint exceeds 2x the L2 (every instance reads from HBM), and the
    replay is bracketed by CUDA events — host launch latency is not in the number.  make_set() -> one set of inputs;
    run(set) runs the op (fwd, or fwd+bwd).  Returns (median ms per op instance, min ms, R)."""
    R = max(2, min(48, -(-2 * L2_BYTES // max(set_bytes, 1))))
    sets = [make_set() for _ in range(R)]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for s in sets[:2]:
            run(s)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for s in sets:
            run(s)
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) / R)
    del g
    return statistics.median(ts), min(ts), R
