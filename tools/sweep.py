"""Kernel-level sweep on one B200: achieved algorithmic GB/s of each HBM-bound kernel vs the measured copy
peak (MEASURED_PEAKS.json).  Algorithmic bytes per SURVEY.md §8d.  Prints one JSON line per point.

    python tools/sweep.py [--quick] [--dtype f32|bf16]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from datamining_recblr_b200 import ops  # noqa: E402
from datamining_recblr_b200.timing import summarize, time_cuda  # noqa: E402


def peak_gbs():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    dt = torch.float32 if args.dtype == "f32" else torch.bfloat16
    es = 4 if dt == torch.float32 else 2
    peak, src = peak_gbs()
    dev = "cuda"
    torch.manual_seed(2020)
    shapes = [(256, L, D) for L in (50, 200, 1024, 4096) for D in (64, 128, 256)] + [(2048, 200, 128), (2048, 50, 128)]
    if args.quick:
        shapes = [(256, 1024, 128), (2048, 200, 128)]

    def emit(**kw):
        kw["peak_gbs"], kw["peak_src"] = peak, src
        kw["frac"] = kw["gbs"] / peak
        print(json.dumps(kw), flush=True)

    for (B, T, C) in shapes:
        E = B * T * C
        xp = torch.randn(B, T, C, device=dev, dtype=dt)
        ri = torch.randn(B, T, 2 * C, device=dev, dtype=dt)
        r, i = ri.chunk(2, -1)
        lam = torch.linspace(-2.2, -6.9, C, device=dev)
        g = torch.randn(B, T, C, device=dev, dtype=dt)
        z = torch.randn(B, T, C, device=dev, dtype=dt)
        xpg, rig, lamg = xp.clone().requires_grad_(), ri.clone().requires_grad_(), lam.clone().requires_grad_()
        # gated scan fwd (4 units) / bwd (8 units)
        t = summarize(time_cuda(lambda: ops.gated_scan(xp, r, i, lam), iters=args.iters))
        emit(kernel="gated_scan_fwd", B=B, T=T, C=C, dtype=args.dtype, ms=t["median_ms"], min_ms=t["min_ms"],
             gbs=4 * E * es / t["median_ms"] / 1e6)
        rg, ig = rig.chunk(2, -1)
        h = ops.gated_scan(xpg, rg, ig, lamg)
        t = summarize(time_cuda(lambda: torch.autograd.grad(h, (xpg, rig, lamg), g, retain_graph=True),
                                iters=args.iters))
        emit(kernel="gated_scan_bwd", B=B, T=T, C=C, dtype=args.dtype, ms=t["median_ms"], min_ms=t["min_ms"],
             gbs=8 * E * es / t["median_ms"] / 1e6)
        # fused z-gate variants: fwd reads 4 writes 2 (6), bwd reads 6 writes 4 (10)
        t = summarize(time_cuda(lambda: ops.gated_scan(xp, r, i, lam, None, z), iters=args.iters))
        emit(kernel="gated_scan_z_fwd", B=B, T=T, C=C, dtype=args.dtype, ms=t["median_ms"], min_ms=t["min_ms"],
             gbs=6 * E * es / t["median_ms"] / 1e6)
        zg = z.clone().requires_grad_()
        y = ops.gated_scan(xpg, rg, ig, lamg, None, zg)
        t = summarize(time_cuda(lambda: torch.autograd.grad(y, (xpg, rig, lamg, zg), g, retain_graph=True),
                                iters=args.iters))
        emit(kernel="gated_scan_z_bwd", B=B, T=T, C=C, dtype=args.dtype, ms=t["median_ms"], min_ms=t["min_ms"],
             gbs=10 * E * es / t["median_ms"] / 1e6)
        del h, y
        # conv fwd (2 units) / bwd (3 units: x, dy -> dx)
        w = torch.randn(C, 4, device=dev) * 0.5
        bias = torch.randn(C, device=dev) * 0.5
        t = summarize(time_cuda(lambda: ops.causal_conv1d_channel_last(xp, w, bias, True), iters=args.iters))
        emit(kernel="conv_fwd", B=B, T=T, C=C, dtype=args.dtype, ms=t["median_ms"], min_ms=t["min_ms"],
             gbs=2 * E * es / t["median_ms"] / 1e6)
        wg, bg = w.clone().requires_grad_(), bias.clone().requires_grad_()
        yc = ops.causal_conv1d_channel_last(xpg, wg, bg, True)
        t = summarize(time_cuda(lambda: torch.autograd.grad(yc, (xpg, wg, bg), g, retain_graph=True), iters=args.iters))
        emit(kernel="conv_bwd", B=B, T=T, C=C, dtype=args.dtype, ms=t["median_ms"], min_ms=t["min_ms"],
             gbs=3 * E * es / t["median_ms"] / 1e6)
        del yc
        if dt == torch.float32:
            # S0 raw scan on [B, C, T]: fwd 3 units, bwd 5 units
            a = torch.rand(B, C, T, device=dev) * 0.5 + 0.5
            b = torch.randn(B, C, T, device=dev)
            t = summarize(time_cuda(lambda: ops.parallel_scan(a, b), iters=args.iters))
            emit(kernel="scan_bct_fwd", B=B, T=T, C=C, dtype="f32", ms=t["median_ms"], min_ms=t["min_ms"],
                 gbs=3 * E * 4 / t["median_ms"] / 1e6)
            ag, bg2 = a.clone().requires_grad_(), b.clone().requires_grad_()
            hh = ops.parallel_scan(ag, bg2)
            gg = torch.randn(B, C, T, device=dev)
            t = summarize(time_cuda(lambda: torch.autograd.grad(hh, (ag, bg2), gg, retain_graph=True), iters=args.iters))
            emit(kernel="scan_bct_bwd", B=B, T=T, C=C, dtype="f32", ms=t["median_ms"], min_ms=t["min_ms"],
                 gbs=5 * E * 4 / t["median_ms"] / 1e6)
            del hh
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
