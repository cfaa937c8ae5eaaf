"""Runs one op a few times at a large shape (for `ncu -k regex:<kernel> -s 2 -c 1 python tools/prof_one.py <op>`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from datamining_recblr_b200 import ops  # noqa: E402

op = sys.argv[1] if len(sys.argv) > 1 else "gscan"
dt = torch.bfloat16 if "bf16" in sys.argv else torch.float32
B, T, C = 2048, 200, 128
torch.manual_seed(0)
xp = torch.randn(B, T, C, device="cuda", dtype=dt, requires_grad=True)
ri = torch.randn(B, T, 2 * C, device="cuda", dtype=dt, requires_grad=True)
z = torch.randn(B, T, C, device="cuda", dtype=dt, requires_grad=True)
lam = torch.linspace(-2.2, -6.9, C, device="cuda", requires_grad=True)
h0 = torch.randn(C, device="cuda", requires_grad=True)
g = torch.randn(B, T, C, device="cuda", dtype=dt)
w = torch.randn(C, 4, device="cuda", requires_grad=True)
bias = torch.randn(C, device="cuda", requires_grad=True)
if op == "cefwd":   # fused CE forward statistics (tcgen05): ncu -k regex:fullsort_kernel
    Bq, N, D = 8192, 300_000, 128
    gq = torch.Generator(device="cuda").manual_seed(0)
    q = torch.randn(Bq, D, device="cuda", generator=gq).to(torch.bfloat16)
    e = (torch.randn(N, D, device="cuda", generator=gq) * 0.05).to(torch.bfloat16)
    pos = torch.randint(0, N, (Bq,), device="cuda", generator=gq)
    for _ in range(4):
        ops.fullsort_ce_stats(q, e, pos)
    torch.cuda.synchronize()
    print("done")
    sys.exit(0)
for _ in range(4):
    if op == "gscan":
        r, i = ri.chunk(2, -1)
        y = ops.gated_scan(xp, r, i, lam, h0=h0, z=z if "z" in sys.argv else None)
        y.backward(g)
    elif op == "conv":
        xz = torch.cat([xp, z], -1)
        y = ops.causal_conv1d_channel_last(xz.chunk(2, -1)[0], w, bias, silu=True)
        y.backward(g)
torch.cuda.synchronize()
print("done")
