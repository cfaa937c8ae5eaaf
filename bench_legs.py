"""Rank-0, single-GPU legs of bench.py's default (north-star) line:

  sweep_leg()   BASELINE.json configs[2]: the S1 fused gate+scan (RecBLR.py:197-200) forward + backward at the model shape
                2 048 x 200 x 128 and the sweep corner 256 x 4096 x 256, fp32 and bf16 I/O, as achieved ALGORITHMIC GB/s
                (12 * E * s bytes) against the measured HBM copy peak.  CUDA-graph replay over rotating input
                sets larger than 2x L2 (datamining_recblr_b200/timing.time_graph).
  triton_leg()  "the kernel to beat" (SURVEY Appendix C / BASELINE.md §3): the reference's own GatedRecurrentLayer
                (literal 200 -> 256 left pad, F.conv1d fallback, separate gate ops, two transposes, Triton
                `parallel_scan`) and its bare `parallel_scan`, timed on the same GPU against this repo's layer / scan at
                the same shapes.  Needs oracle/_ref/reference_py.tar.gz (oracle/build_ref.py); reports why when absent.
"""
import statistics

import torch


def _events(fn, warmup, iters, flush):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    return statistics.median(ts), min(ts)


def sweep_leg(peak_gbs, shapes=((2048, 200, 128), (256, 4096, 256)), dtypes=("f32", "bf16")):
    from datamining_recblr_b200 import ops
    from datamining_recblr_b200.timing import time_graph
    dev = "cuda"
    out = []
    for (B, T, C) in shapes:
        for dname in dtypes:
            dt = torch.float32 if dname == "f32" else torch.bfloat16
            es = 4 if dname == "f32" else 2
            E = B * T * C * es
            lam0 = torch.linspace(-2.2, -6.9, C, device=dev)

            def gset():
                return dict(xp=torch.randn(B, T, C, device=dev, dtype=dt).requires_grad_(),
                            ri=(torch.randn(B, T, 2 * C, device=dev, dtype=dt)).requires_grad_(),
                            lam=lam0.clone().requires_grad_(), g=torch.randn(B, T, C, device=dev, dtype=dt))

            def g_fb(s):
                y = ops.gated_scan_packed(s["xp"], s["ri"], s["lam"])
                s["xp"].grad = s["ri"].grad = s["lam"].grad = None
                y.backward(s["g"])

            med, mn, R = time_graph(gset, g_fb, 16 * E, iters=5)
            gbs = 12 * E / med / 1e6
            out.append(dict(op="gated_scan fwd+bwd (S1: 12*E*s algorithmic bytes)", B=B, T=T, C=C, dtype=dname, ms=med,
                            ms_min=mn, gbs=gbs, frac=gbs / peak_gbs, seq_tokens_per_s=B * T / (med * 1e-3),
                            rotating_sets=R))
            torch.cuda.empty_cache()
    return out


def triton_leg(B=2048, L=200, d_model=64, iters=5):
    from oracle import build_ref
    if not build_ref.available():
        return {"unavailable": "oracle/_ref/reference_py.tar.gz not built (python -m oracle.build_ref where /root/reference exists)"}
    from datamining_recblr_b200 import ops
    from datamining_recblr_b200.recblr import GatedRecurrentLayer
    from datamining_recblr_b200.timing import flush_l2
    try:
        mod, ps = build_ref.load_gpu_reference()
    except Exception as exc:  # the reference needs its pinned triton/torch to import: report, do not hide
        return {"unavailable": f"reference import failed: {type(exc).__name__}: {exc}"}
    dev = "cuda"
    torch.manual_seed(0)
    ref = mod.GatedRecurrentLayer(d_model=d_model, expansion_factor=2, kernel_size=4).to(dev)
    ours = GatedRecurrentLayer(d_model=d_model, expansion_factor=2, kernel_size=4).to(dev)
    ours.load_state_dict(ref.state_dict())
    x = torch.randn(B, L, d_model, device=dev)
    gy = torch.randn(B, L, d_model, device=dev)

    def layer_step(m, xin, g, amp):
        def run():
            xi = xin.detach().requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                y = m(xi)
            y.backward(g.to(y.dtype))
            m.zero_grad(set_to_none=True)
        return run

    res = {"shape": f"GatedRecurrentLayer fwd+bwd, B={B} L={L} d_model={d_model} (C={2 * d_model}); reference pads {L}->"
                    f"{2 ** ((L - 1).bit_length())} and runs its Triton scan"}
    try:
        t_ref, _ = _events(layer_step(ref, x, gy, False), 2, iters, flush_l2)
    except Exception as exc:
        return {"unavailable": f"reference layer failed to run on this GPU: {type(exc).__name__}: {exc}"}
    t_f32, _ = _events(layer_step(ours, x, gy, False), 3, iters, flush_l2)
    t_bf16, _ = _events(layer_step(ours, x, gy, True), 3, iters, flush_l2)
    res.update(reference_layer_ms=t_ref, ours_layer_f32_ms=t_f32, ours_layer_bf16_ms=t_bf16,
               layer_ratio=t_ref / t_f32, layer_ratio_bf16=t_ref / t_bf16)

    # inference (no_grad): the reference layer's forward vs this repo's, whose conv + gates GEMM + recurrence + z-gate is ONE
    # tcgen05 kernel at this shape (ops.bdlru_core_fused, bf16 autocast)
    def layer_infer(m, amp):
        def run():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                m(x)
        return run
    ours.eval(), ref.eval()
    i_ref, _ = _events(layer_infer(ref, False), 2, iters, flush_l2)
    i_ours, _ = _events(layer_infer(ours, True), 3, iters, flush_l2)
    ours.train(), ref.train()
    res.update(reference_layer_infer_ms=i_ref, ours_layer_infer_bf16_ms=i_ours, layer_infer_ratio=i_ref / i_ours)
    # bare scan op at the reference's own contract: contiguous fp32 [B, C, Tp], power-of-two Tp
    C, Tp = 2 * d_model, 2 ** ((L - 1).bit_length())
    a = (torch.rand(B, C, Tp, device=dev) * 0.5 + 0.5)
    b = torch.randn(B, C, Tp, device=dev)
    g = torch.randn(B, C, Tp, device=dev)

    def scan_step(fn):
        def run():
            ai, bi = a.detach().requires_grad_(True), b.detach().requires_grad_(True)
            fn(ai, bi).backward(g)
        return run
    s_ref, _ = _events(scan_step(ps.parallel_scan), 2, iters, flush_l2)
    s_ours, _ = _events(scan_step(ops.parallel_scan), 3, iters, flush_l2)
    res.update(reference_scan_ms=s_ref, ours_scan_ms=s_ours, scan_ratio=s_ref / s_ours,
               scan_shape=f"parallel_scan fwd+bwd fp32 [B={B}, C={C}, T={Tp}]")
    # what the layer's scan stage costs each side: reference = gate ops + 2 transposes + Triton scan on Tp steps;
    # ours = one fused gate+scan kernel pair on the L real steps
    xp = torch.randn(B, L, C, device=dev, requires_grad=True)
    ri = torch.randn(B, L, 2 * C, device=dev, requires_grad=True)
    lam = ref.Lambda.detach().clone().requires_grad_(True)
    gh = torch.randn(B, L, C, device=dev)

    def ours_stage():
        xp.grad = ri.grad = lam.grad = None
        ops.gated_scan_packed(xp, ri, lam).backward(gh)

    def ref_stage():
        xp.grad = ri.grad = lam.grad = None
        import torch.nn.functional as F
        pad = Tp - L
        xpp, rip = F.pad(xp, (0, 0, pad, 0)), F.pad(ri, (0, 0, pad, 0))
        r, i = rip.chunk(2, dim=-1)
        alpha = torch.exp(-F.softplus(lam) * torch.sigmoid(r))
        beta = torch.sqrt(1 - alpha.pow(2) + 1e-8) * torch.sigmoid(i) * xpp
        h = ps.parallel_scan(alpha.mT.contiguous(), beta.mT.contiguous()).mT[:, pad:]
        h.backward(gh)
    g_ref, _ = _events(ref_stage, 2, iters, flush_l2)
    g_ours, _ = _events(ours_stage, 3, iters, flush_l2)
    res.update(reference_gate_scan_stage_ms=g_ref, ours_gate_scan_stage_ms=g_ours, gate_scan_stage_ratio=g_ref / g_ours)
    return res
