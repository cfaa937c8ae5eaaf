"""Inference core of the BD-LRU layer: ONE fused tcgen05 kernel (ops.bdlru_core_fused) vs the separate kernels + cuBLAS
(ops.bdlru_block under no_grad: conv -> gates GEMM -> gate+scan+z).  Graph replay over rotating input sets > 2x L2."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from datamining_recblr_b200 import ops  # noqa: E402
from datamining_recblr_b200.timing import time_graph  # noqa: E402

C = 128
dev = "cuda"
torch.manual_seed(0)
conv_w, conv_b = torch.randn(C, 4, device=dev) * 0.5, torch.randn(C, device=dev) * 0.5
gates_w = (torch.randn(2 * C, C, device=dev) * 0.15)
gates_wb = gates_w.to(torch.bfloat16)
gates_b = torch.randn(2 * C, device=dev) * 0.5
lam = torch.linspace(-2.2, -6.9, C, device=dev)
h0 = torch.randn(C, device=dev)
for (B, T) in [(4096, 50), (2048, 200), (4096, 200), (16384, 200), (512, 1024)]:
    E = B * T * C * 2

    def mk():
        return torch.randn(B, T, 2 * C, device=dev).to(torch.bfloat16)

    def fused(xz):
        ops.bdlru_core_fused(xz, conv_w, conv_b, gates_wb, gates_b, lam, h0=h0)

    def separate(xz):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            ops.bdlru_block(xz, conv_w, conv_b, gates_w, gates_b, lam, h0=h0)

    tf, _, R = time_graph(mk, fused, 3 * E, iters=5)
    ts, _, _ = time_graph(mk, separate, 8 * E, iters=5)
    print(f"B={B} T={T} C={C}: fused {tf:.4f} ms ({3 * E / tf / 1e6:.0f} GB/s of its 3-unit traffic, "
          f"{B * T * C / tf / 1e6:.1f} G elem/s)   separate kernels {ts:.4f} ms   speed-up {ts / tf:.2f}x   (R={R})")

# ---- the whole first half of RecurrentLayer (in-proj .. residual LayerNorm): ONE kernel vs the module's separate path
from datamining_recblr_b200.recblr import RecurrentLayer  # noqa: E402

D = 64
layer = RecurrentLayer(d_model=D, d_conv=4, expand=2, dropout=0.0, num_layers=1, bd_lru_only=False, disable_conv1d=False,
                       disable_ffn=True).cuda().eval()
for (B, T) in [(4096, 50), (2048, 200), (4096, 200), (16384, 200)]:
    E = B * T * D * 2

    def mkx():
        return torch.randn(B, T, D, device=dev).to(torch.bfloat16)

    def run_layer(fused_layer, fused_core):
        def run(x):
            layer.fused_layer = fused_layer
            layer.behavior_modeling.fused_core = fused_core
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                layer(x)
        return run

    t_full, _, R = time_graph(mkx, run_layer(True, True), 4 * E, iters=5)
    t_core, _, _ = time_graph(mkx, run_layer(False, True), 16 * E, iters=5)
    t_sep, _, _ = time_graph(mkx, run_layer(False, False), 30 * E, iters=5)
    print(f"layer first half B={B} T={T} D={D}: one kernel {t_full:.4f} ms ({2 * E / t_full / 1e6:.0f} GB/s of its 2 D-wide units, "
          f"{B * T * 128 / t_full / 1e6:.1f} G elem/s) | cuBLAS proj + fused core + add_ln {t_core:.4f} ms | all separate "
          f"{t_sep:.4f} ms | speed-up {t_sep / t_full:.2f}x / {t_core / t_full:.2f}x")
