"""Throughput of the fused CE forward / backward kernels at a large shape (tuning aid)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from datamining_recblr_b200 import ops  # noqa: E402

B, N, D = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000, 128
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.randn(B, D, device="cuda", generator=g).to(torch.bfloat16)
e = (torch.randn(N, D, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
pos = torch.randint(0, N, (B,), device="cuda", generator=g)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


m, s, pl = ops.fullsort_ce_stats(q, e, pos)
lse = m + torch.log(s)
t_f = timeit(lambda: ops.fullsort_ce_stats(q, e, pos))
t_b = timeit(lambda: ops.fullsort_ce_grads(q, e, pos, lse, 1.0 / B))
fl = 2.0 * B * N * D
print(f"B={B} N={N} D={D}: ce fwd {t_f:.3f} ms = {fl / t_f / 1e9:.0f} TFLOP/s;  ce bwd (dQ+dE) {t_b:.3f} ms = "
      f"{2 * fl / t_b / 1e9:.0f} TFLOP/s credited (4 GEMM-passes executed: {4 * fl / t_b / 1e9:.0f} TFLOP/s)")

# split of the backward: dQ pass only / dE pass only (through the C ABI directly)
from datamining_recblr_b200 import _lib as L  # noqa: E402
lib = L.load()
nws = lib.bdlru_fullsort_ce_workspace_bytes(B, N, D)
ws = torch.empty(max(nws, 16), dtype=torch.uint8, device="cuda")
dQ = torch.empty(B, D, device="cuda")
dE = torch.empty(N, D, device="cuda")
st = L.stream_ptr(q)
lsef = lse.float().contiguous()


def run(dq, de):
    L.check(lib.bdlru_fullsort_ce_bwd(L.ptr(q), L.ptr(e), L.ptr(pos), L.ptr(lsef), 1.0 / B, None, B, N, D, 0, L.ptr(dq),
                                      L.ptr(de), L.ptr(ws), nws, st))


t_q = timeit(lambda: run(dQ, None))
t_e = timeit(lambda: run(None, dE))
print(f"  dQ pass {t_q:.3f} ms ({2 * fl / t_q / 1e9:.0f} TFLOP/s executed), dE pass {t_e:.3f} ms ({2 * fl / t_e / 1e9:.0f} TFLOP/s executed)")

# fused forward of the training step: reference maximum (sampled / exact) + one exponential pass (denominator + dQ)
for stride in (16, 1):
    t_m = timeit(lambda: ops.fullsort_rowmax(q, e, stride))
    print(f"  rowmax stride {stride}: {t_m:.3f} ms ({fl / stride / t_m / 1e9:.0f} TFLOP/s executed)")
ref = ops.fullsort_rowmax(q, e, 16)
t_fd = timeit(lambda: ops.fullsort_ce_fwd_dq(q, e, ref))
print(f"  fused fwd+dQ pass {t_fd:.3f} ms ({2 * fl / t_fd / 1e9:.0f} TFLOP/s, all credited)")
print(f"  training CE total: fused {t_fd + t_e:.3f} ms + reference  vs  stats fwd + dQ + dE {t_f + t_q + t_e:.3f} ms")
