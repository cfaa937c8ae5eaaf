"""Every kernel of libbdlru.so once at small shapes (for `compute-sanitizer --tool memcheck|racecheck python tools/sanitize_smoke.py`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from datamining_recblr_b200 import ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
for dt in (torch.float32, torch.bfloat16):
    B, T, C = 3, 37, 64
    xp = torch.randn(B, T, C, device=dev, dtype=dt, requires_grad=True)
    ri = torch.randn(B, T, 2 * C, device=dev, dtype=dt, requires_grad=True)
    z = torch.randn(B, T, C, device=dev, dtype=dt, requires_grad=True)
    lam = torch.linspace(-2.2, -6.9, C, device=dev, requires_grad=True)
    h0 = torch.randn(C, device=dev, requires_grad=True)
    ops.gated_scan_packed(xp, ri, lam, h0, z=z).sum().backward()
    r, i = ri.chunk(2, -1)
    ops.gated_scan(xp, r, i, lam, torch.randn(B, C, device=dev, requires_grad=True)).sum().backward()
    a = torch.rand(B, T, C, device=dev, dtype=dt, requires_grad=True)
    ops.scan_channel_last(a, xp, h0).sum().backward()
    w = torch.randn(C, 4, device=dev, requires_grad=True)
    b = torch.randn(C, device=dev, requires_grad=True)
    ops.causal_conv1d_channel_last(xp, w, b, silu=True).sum().backward()
    table = torch.randn(50, C, device=dev, dtype=dt, requires_grad=True)
    ids = torch.randint(0, 50, (B, T), device=dev)
    g, bt = torch.ones(C, device=dev, requires_grad=True), torch.zeros(C, device=dev, requires_grad=True)
    y = ops.embed_layernorm(ids, table, g, bt, dropout_p=0.1, seed=3, padding_idx=0)
    ops.add_dropout_layernorm(y, y.detach(), g, bt, dropout_p=0.1, seed=5).sum().backward()
    ops.colsum(torch.randn(101, 64, device=dev, dtype=dt))
ga = torch.rand(2, 5, 33, device=dev, requires_grad=True)
gb = torch.randn(2, 5, 33, device=dev, requires_grad=True)
ops.parallel_scan(ga, gb).sum().backward()
for (Bq, N, D) in ((130, 700, 64), (5, 300, 128), (140, 200, 256)):
    q = torch.randn(Bq, D, device=dev, requires_grad=True)
    e = (torch.randn(N, D, device=dev) * 0.3).requires_grad_()
    pos = torch.randint(0, N, (Bq,), device=dev)
    s, i = ops.fullsort_topk(q, e, 10, mask_id=0)
    ops.topk_merge(torch.cat([s, s], 1), torch.cat([i, i + N], 1), 10)
    ops.fullsort_cross_entropy(q, e, pos).backward()
torch.cuda.synchronize()
print("sanitize_smoke done")
