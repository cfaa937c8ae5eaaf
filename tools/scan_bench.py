"""Times the fused gate+scan (+z) forward / backward at one shape (graph replay over rotating sets), for ncu captures.
    python tools/scan_bench.py [B T C dtype z]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from datamining_recblr_b200 import ops  # noqa: E402
from datamining_recblr_b200.timing import time_graph  # noqa: E402

B, T, C = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (2048, 200, 128)
dt = torch.bfloat16 if (len(sys.argv) <= 4 or sys.argv[4] == "bf16") else torch.float32
z = len(sys.argv) > 5 and sys.argv[5] == "z"
es = 2 if dt == torch.bfloat16 else 4
E = B * T * C * es
dev = "cuda"
lam0 = torch.linspace(-2.2, -6.9, C, device=dev)


def gset():
    d = dict(xp=torch.randn(B, T, C, device=dev, dtype=dt).requires_grad_(),
             ri=torch.randn(B, T, 2 * C, device=dev, dtype=dt).requires_grad_(),
             lam=lam0.clone().requires_grad_(), g=torch.randn(B, T, C, device=dev, dtype=dt))
    if z:
        d["z"] = torch.randn(B, T, C, device=dev, dtype=dt).requires_grad_()
    return d


def fwd(s):
    with torch.no_grad():
        ops.gated_scan_packed(s["xp"], s["ri"], s["lam"], z=s.get("z"))


def fb(s):
    y = ops.gated_scan_packed(s["xp"], s["ri"], s["lam"], z=s.get("z"))
    s["xp"].grad = s["ri"].grad = s["lam"].grad = None
    if z:
        s["z"].grad = None
    y.backward(s["g"])


fu, bu = (6, 10) if z else (4, 8)
mf, _, R = time_graph(gset, fwd, (fu + 2) * E, iters=5)
mfb, _, _ = time_graph(gset, fb, (fu + bu + 4) * E, iters=5)
peak = 6543.1
print(f"{B}x{T}x{C} {dt} z={z}: fwd {mf:.4f} ms = {fu * E / mf / 1e6:.0f} GB/s ({fu * E / mf / 1e6 / peak:.2f});  fwd+bwd {mfb:.4f} ms = "
      f"{(fu + bu) * E / mfb / 1e6:.0f} GB/s ({(fu + bu) * E / mfb / 1e6 / peak:.2f});  bwd alone ~{mfb - mf:.4f} ms = "
      f"{bu * E / (mfb - mf) / 1e6:.0f} GB/s ({bu * E / (mfb - mf) / 1e6 / peak:.2f})")
