"""BASELINE.json configs[2] — synthetic BD-LRU scan sweep on one B200: B = 256, L in {50, 200, 1024, 4096},
D in {64, 128, 256} (+ the model shapes), forward and backward of the S1 fused gate+scan (+ z-gate), the S0 raw scan and
the causal conv, as achieved ALGORITHMIC GB/s against the measured HBM copy peak (MEASURED_PEAKS.json).

Timing: each op is captured into a CUDA graph that runs it on R rotating input sets whose total footprint exceeds 2x the
L2 (so every instance reads from HBM, no explicit flush inside the timed region) and the graph replay is bracketed by
CUDA events — host launch latency is not in the number.  Algorithmic bytes per SURVEY.md §8d / DESIGN.md §2.

    python tools/sweep.py [--quick] [--dtype f32|bf16] > gpurun_out/sweep.jsonl
"""
import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from datamining_recblr_b200 import ops  # noqa: E402

L2_BYTES = 126 << 20


def peak_gbs():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def time_graph(make_set, run, set_bytes, iters=7):
    from datamining_recblr_b200.timing import time_graph as tg
    med, _, R = tg(make_set, run, set_bytes, iters)
    return med, R


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--dtype", default="f32")
    args = ap.parse_args()
    dt = torch.float32 if args.dtype == "f32" else torch.bfloat16
    es = 4 if dt == torch.float32 else 2
    peak, src = peak_gbs()
    dev = "cuda"
    torch.manual_seed(2020)
    shapes = [(256, L, D) for L in (50, 200, 1024, 4096) for D in (64, 128, 256)] + [(2048, 200, 128), (2048, 50, 128)]
    if args.quick:
        shapes = [(256, 1024, 128), (256, 4096, 256), (2048, 200, 128)]

    def emit(kernel, B, T, C, units, ms, R):
        gbs = units * B * T * C * es / ms / 1e6
        print(json.dumps(dict(kernel=kernel, B=B, T=T, C=C, dtype=args.dtype, ms=ms, units=units, gbs=gbs, peak_gbs=peak,
                              peak_src=src, frac=gbs / peak, rotating_sets=R, tokens_per_s=B * T / (ms * 1e-3))), flush=True)

    for (B, T, C) in shapes:
        E = B * T * C * es
        lam0 = torch.linspace(-2.2, -6.9, C, device=dev)

        def gset(z):
            d = dict(xp=torch.randn(B, T, C, device=dev, dtype=dt).requires_grad_(),
                     ri=torch.randn(B, T, 2 * C, device=dev, dtype=dt).requires_grad_(),
                     lam=lam0.clone().requires_grad_(), g=torch.randn(B, T, C, device=dev, dtype=dt))
            if z:
                d["z"] = torch.randn(B, T, C, device=dev, dtype=dt).requires_grad_()
            return d

        def g_fwd(s):
            with torch.no_grad():
                return ops.gated_scan_packed(s["xp"], s["ri"], s["lam"], z=s.get("z"))

        def g_fb(s):
            y = ops.gated_scan_packed(s["xp"], s["ri"], s["lam"], z=s.get("z"))
            s["xp"].grad = s["ri"].grad = s["lam"].grad = None
            if "z" in s:
                s["z"].grad = None
            y.backward(s["g"])

        for z in (False, True):
            tag = "gated_scan_z" if z else "gated_scan"
            fu, bu = (6, 10) if z else (4, 8)
            ms_f, R = time_graph(lambda: gset(z), g_fwd, (fu + 2) * E)
            emit(tag + "_fwd", B, T, C, fu, ms_f, R)
            ms_fb, R = time_graph(lambda: gset(z), g_fb, (fu + bu + 4) * E)
            emit(tag + "_fwd_bwd", B, T, C, fu + bu, ms_fb, R)

        def cset():
            return dict(x=torch.randn(B, T, C, device=dev, dtype=dt).requires_grad_(),
                        w=(torch.randn(C, 4, device=dev) * 0.5).requires_grad_(),
                        b=(torch.randn(C, device=dev) * 0.5).requires_grad_(), g=torch.randn(B, T, C, device=dev, dtype=dt))

        def c_fwd(s):
            with torch.no_grad():
                return ops.causal_conv1d_channel_last(s["x"], s["w"], s["b"], silu=True)

        def c_fb(s):
            y = ops.causal_conv1d_channel_last(s["x"], s["w"], s["b"], silu=True)
            s["x"].grad = s["w"].grad = s["b"].grad = None
            y.backward(s["g"])

        ms_f, R = time_graph(cset, c_fwd, 3 * E)
        emit("conv_fwd", B, T, C, 2, ms_f, R)
        ms_fb, R = time_graph(cset, c_fb, 6 * E)
        emit("conv_fwd_bwd", B, T, C, 6, ms_fb, R)

        if dt == torch.float32:
            def sset():
                return dict(a=torch.rand(B, C, T, device=dev).mul_(0.5).add_(0.5).requires_grad_(),
                            b=torch.randn(B, C, T, device=dev).requires_grad_(), g=torch.randn(B, C, T, device=dev))

            def s_fwd(s):
                with torch.no_grad():
                    return ops.parallel_scan(s["a"], s["b"])

            def s_fb(s):
                y = ops.parallel_scan(s["a"], s["b"])
                s["a"].grad = s["b"].grad = None
                y.backward(s["g"])

            ms_f, R = time_graph(sset, s_fwd, 4 * E)
            emit("scan_bct_fwd", B, T, C, 3, ms_f, R)
            ms_fb, R = time_graph(sset, s_fb, 8 * E)
            emit("scan_bct_fwd_bwd", B, T, C, 8, ms_fb, R)


if __name__ == "__main__":
    main()
