import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from datamining_recblr_b200 import ops
C = 128
for (B, T, use_conv) in [(296, 50, False), (297, 50, False), (300, 50, False), (300, 50, True), (600, 50, True), (600, 50, False)]:
    for rep in range(3):
        g = torch.Generator(device="cuda").manual_seed(B * 1000 + T)
        rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
        xz = rn(B, T, 2 * C).to(torch.bfloat16)
        conv_w, conv_b = rn(C, 4) * 0.5, rn(C) * 0.5
        gates_w = (rn(2 * C, C) * 0.15).to(torch.bfloat16).float()
        gates_b = rn(2 * C) * 0.5
        lam = torch.linspace(-2.2, -6.9, C, device="cuda") + 0.3 * rn(C)
        h0 = rn(C)
        y = ops.bdlru_core_fused(xz, conv_w, conv_b, gates_w, gates_b, lam, h0=h0, use_conv=use_conv)
        with torch.no_grad():
            ys = ops.bdlru_block(xz, conv_w, conv_b, gates_w, gates_b, lam, h0=h0, use_conv=use_conv)
        bad = ~torch.isfinite(y.float())
        err = (y.float() - ys.float()).abs()
        err[bad] = 0
        idx = bad.nonzero()
        print(B, T, use_conv, rep, "nonfinite", int(bad.sum()), "first", idx[:3].tolist(), "max err", float(err.max()),
              "rows with err>0.1:", sorted(set((err > 0.1).nonzero()[:, 0].tolist()))[:10])
