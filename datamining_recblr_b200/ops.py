"""Differentiable PyTorch operators over the C ABI (host-side plumbing only: shapes, strides, autograd).

    parallel_scan(gates, tokens)                       S0, [B, C, T]      parallel_scan.py:117-118
    gated_scan(xp, r, i, Lambda, h0=None, z=None)      S1, [B, T, C]      RecBLR.py:197-200 (+206 with z)
    scan_channel_last(a, b, h0=None)                   raw scan, [B, T, C]
    causal_conv1d_channel_last(x, weight, bias, silu)  [B, T, C]          RecBLR.py:185 / 188-193
"""
import ctypes

import torch

from . import _lib as L


class _ScanBCT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gates, tokens):
        L.require_cuda(gates, tokens)
        B, C, T = gates.shape
        # same contract as the reference op (parallel_scan.py:87-89)
        assert tokens.shape == (B, C, T)
        assert gates.is_contiguous()
        assert tokens.is_contiguous()
        assert gates.dtype == torch.float32 and tokens.dtype == torch.float32
        states = torch.empty_like(tokens)
        L.check(L.load().bdlru_scan_fwd(L.ptr(gates), L.ptr(tokens), L.ptr(states), B * C, T, L.stream_ptr(gates)))
        ctx.save_for_backward(states, gates)
        return states

    @staticmethod
    def backward(ctx, grad_output):
        states, gates = ctx.saved_tensors
        B, C, T = gates.shape
        grad_output = grad_output.contiguous()
        d_gates = torch.empty_like(gates)
        d_tokens = torch.empty_like(gates)
        L.check(L.load().bdlru_scan_bwd(L.ptr(gates), L.ptr(states), L.ptr(grad_output), L.ptr(d_gates),
                                        L.ptr(d_tokens), B * C, T, L.stream_ptr(gates)))
        return d_gates, d_tokens


def parallel_scan(gates, tokens):
    """Drop-in for the reference's parallel_scan (parallel_scan.py:117): h_t = gates_t*h_{t-1} + tokens_t over
    the last axis of contiguous fp32 [B, C, T]; any T (no power-of-two restriction)."""
    return _ScanBCT.apply(gates, tokens)


def _workspace(device, nbytes):
    """Scratch for ONE kernel call, owned by the caching allocator (or, under capture, by the CUDA graph's private pool):
    its lifetime follows stream order like any other tensor, so a captured graph can never be left pointing at a buffer
    that a later, larger eager call replaced (a cached grow-on-demand buffer had exactly that hazard), and two streams
    never share scratch."""
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


class _GatedScan(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xp, r, i, Lambda, h0, z):
        L.require_cuda(xp, r, i, Lambda, h0, z)
        B, T, C = xp.shape
        assert r.shape == (B, T, C) and i.shape == (B, T, C) and Lambda.shape == (C,)
        dt = L.dtype_tag(xp)
        assert r.dtype == xp.dtype and i.dtype == xp.dtype
        xp, r, i = L.as_cl(xp), L.as_cl(r), L.as_cl(i)
        Lf = Lambda.detach().float().contiguous()
        h0f, h0_bs = None, 0
        if h0 is not None:
            h0f = h0.detach().float().contiguous()
            assert h0f.shape in ((C,), (B, C))
            h0_bs = C if h0f.dim() == 2 else 0
        h = torch.empty((B, T, C), dtype=xp.dtype, device=xp.device)
        y = None
        if z is not None:
            assert z.shape == (B, T, C) and z.dtype == xp.dtype
            z = L.as_cl(z)
            y = torch.empty_like(h)
        L.check(L.load().bdlru_gated_scan_fwd(L.view3(xp), L.view3(r), L.view3(i), L.ptr(Lf), L.ptr(h0f), h0_bs,
                                              L.view3(z), L.view3(h), L.view3(y), B, T, C, dt, L.stream_ptr(xp)))
        ctx.save_for_backward(xp, r, i, Lf, h0f, z, h)
        ctx.h0_bs = h0_bs
        ctx.lambda_dtype = Lambda.dtype
        ctx.h0_dtype = h0.dtype if h0 is not None else None
        if z is not None:
            ctx.mark_non_differentiable(h)
            return y, h
        return h

    @staticmethod
    def backward(ctx, grad, *unused):
        xp, r, i, Lf, h0f, z, h = ctx.saved_tensors
        B, T, C = xp.shape
        dt = L.dtype_tag(xp)
        grad = L.as_cl(grad.to(xp.dtype))
        dxp, dr, di = torch.empty_like(h), torch.empty_like(h), torch.empty_like(h)
        dz = torch.empty_like(h) if z is not None else None
        dLambda = torch.empty(C, dtype=torch.float32, device=xp.device)
        dh0 = torch.empty_like(h0f) if h0f is not None else None
        lib = L.load()
        nws = lib.bdlru_gated_scan_bwd_workspace_bytes(B, T, C)
        ws = _workspace(xp.device, nws)
        L.check(lib.bdlru_gated_scan_bwd(L.view3(xp), L.view3(r), L.view3(i), L.ptr(Lf), L.ptr(h0f), ctx.h0_bs,
                                         L.view3(z), L.view3(h), L.view3(grad), L.view3(dxp), L.view3(dr),
                                         L.view3(di), L.view3(dz), L.ptr(dLambda), L.ptr(dh0), L.ptr(ws), nws,
                                         B, T, C, dt, L.stream_ptr(xp)))
        return (dxp, dr, di, dLambda.to(ctx.lambda_dtype), dh0.to(ctx.h0_dtype) if dh0 is not None else None, dz)


def gated_scan(xp, r, i, Lambda, h0=None, z=None):
    """Fused BD-LRU recurrence on channel-last [B, T, C] views (RecBLR.py:197-200 without transposes/padding):
        a = exp(-softplus(Lambda)*sigmoid(r)); b = sqrt(1-a^2+1e-8)*sigmoid(i)*xp; h_t = a_t*h_{t-1} + b_t, h_{-1}=h0.
    Returns h, or silu(z)*h when z is given (RecBLR.py:206's gate).  fp32 or bf16 I/O, fp32 state."""
    out = _GatedScan.apply(xp, r, i, Lambda, h0, z)
    return out[0] if z is not None else out


class _GatedScanPacked(torch.autograd.Function):
    """Same kernels as _GatedScan with r and i given as the two halves of ONE [B, T, 2C] tensor (the output of the
    `gates` Linear, RecBLR.py:196): the gradient comes back as one [B, T, 2C] tensor written in place by the kernel, so
    autograd does not have to concatenate dr and di."""

    @staticmethod
    def forward(ctx, xp, ri, Lambda, h0, z):
        L.require_cuda(xp, ri, Lambda, h0, z)
        B, T, C = xp.shape
        assert ri.shape == (B, T, 2 * C) and Lambda.shape == (C,) and ri.dtype == xp.dtype
        dt = L.dtype_tag(xp)
        xp, ri = L.as_cl(xp), L.as_cl(ri)
        r, i = ri[..., :C], ri[..., C:]
        if not (L.cl_ok(r) and L.cl_ok(i)):
            ri = ri.contiguous()
            r, i = ri[..., :C], ri[..., C:]
        Lf = Lambda.detach().float().contiguous()
        h0f, h0_bs = None, 0
        if h0 is not None:
            h0f = h0.detach().float().contiguous()
            assert h0f.shape in ((C,), (B, C))
            h0_bs = C if h0f.dim() == 2 else 0
        h = torch.empty((B, T, C), dtype=xp.dtype, device=xp.device)
        y = None
        if z is not None:
            assert z.shape == (B, T, C) and z.dtype == xp.dtype
            z = L.as_cl(z)
            y = torch.empty_like(h)
        L.check(L.load().bdlru_gated_scan_fwd(L.view3(xp), L.view3(r), L.view3(i), L.ptr(Lf), L.ptr(h0f), h0_bs,
                                              L.view3(z), L.view3(h), L.view3(y), B, T, C, dt, L.stream_ptr(xp)))
        ctx.save_for_backward(xp, ri, Lf, h0f, z, h)
        ctx.h0_bs = h0_bs
        ctx.lambda_dtype = Lambda.dtype
        ctx.h0_dtype = h0.dtype if h0 is not None else None
        if z is not None:
            ctx.mark_non_differentiable(h)
            return y, h
        return h

    @staticmethod
    def backward(ctx, grad, *unused):
        xp, ri, Lf, h0f, z, h = ctx.saved_tensors
        B, T, C = xp.shape
        dt = L.dtype_tag(xp)
        r, i = ri[..., :C], ri[..., C:]
        grad = L.as_cl(grad.to(xp.dtype))
        dxp = torch.empty_like(h)
        dri = torch.empty((B, T, 2 * C), dtype=xp.dtype, device=xp.device)
        dz = torch.empty_like(h) if z is not None else None
        dLambda = torch.empty(C, dtype=torch.float32, device=xp.device)
        dh0 = torch.empty_like(h0f) if h0f is not None else None
        lib = L.load()
        nws = lib.bdlru_gated_scan_bwd_workspace_bytes(B, T, C)
        ws = _workspace(xp.device, nws)
        L.check(lib.bdlru_gated_scan_bwd(L.view3(xp), L.view3(r), L.view3(i), L.ptr(Lf), L.ptr(h0f), ctx.h0_bs,
                                         L.view3(z), L.view3(h), L.view3(grad), L.view3(dxp), L.view3(dri[..., :C]),
                                         L.view3(dri[..., C:]), L.view3(dz), L.ptr(dLambda), L.ptr(dh0), L.ptr(ws), nws,
                                         B, T, C, dt, L.stream_ptr(xp)))
        return (dxp, dri, dLambda.to(ctx.lambda_dtype), dh0.to(ctx.h0_dtype) if dh0 is not None else None, dz)


def gated_scan_packed(xp, ri, Lambda, h0=None, z=None):
    """gated_scan with (r, i) = ri.chunk(2, -1) passed as one tensor (see _GatedScanPacked)."""
    out = _GatedScanPacked.apply(xp, ri, Lambda, h0, z)
    return out[0] if z is not None else out


class _PhantomH0(torch.autograd.Function):
    @staticmethod
    def forward(ctx, conv_bias, gates_w, gates_b, Lambda, pad_len):
        L.require_cuda(conv_bias, gates_w, gates_b, Lambda)
        C = Lambda.shape[0]
        assert gates_w.shape == (2 * C, C) and gates_b.shape == (2 * C,) and conv_bias.shape == (C,)
        cb, gw, gb, lam = (t.detach().float().contiguous() for t in (conv_bias, gates_w, gates_b, Lambda))
        h0 = torch.empty(C, dtype=torch.float32, device=lam.device)
        saved = torch.empty(5 * C, dtype=torch.float32, device=lam.device)
        L.check(L.load().bdlru_phantom_h0_fwd(L.ptr(cb), L.ptr(gw), L.ptr(gb), L.ptr(lam), C, int(pad_len), L.ptr(h0),
                                              L.ptr(saved), L.stream_ptr(lam)))
        ctx.save_for_backward(cb, gw, lam, saved)
        ctx.meta = (int(pad_len), conv_bias.dtype, gates_w.dtype, gates_b.dtype, Lambda.dtype)
        return h0

    @staticmethod
    def backward(ctx, dh0):
        cb, gw, lam, saved = ctx.saved_tensors
        pad_len, d0, d1, d2, d3 = ctx.meta
        C = lam.shape[0]
        dcb = torch.empty_like(cb)
        dgw = torch.empty_like(gw)
        dgb = torch.empty(2 * C, dtype=torch.float32, device=lam.device)
        dlam = torch.empty_like(lam)
        L.check(L.load().bdlru_phantom_h0_bwd(L.ptr(cb), L.ptr(gw), L.ptr(lam), L.ptr(saved),
                                              L.ptr(dh0.float().contiguous()), C, pad_len, L.ptr(dcb), L.ptr(dgw), L.ptr(dgb),
                                              L.ptr(dlam), L.stream_ptr(lam)))
        return dcb.to(d0), dgw.to(d1), dgb.to(d2), dlam.to(d3), None


def phantom_h0(conv_bias, gates_w, gates_b, Lambda, pad_len):
    """Initial state equivalent to the reference's `pad_len` left-padded phantom steps (RecBLR.py:177-199), fp32 [C],
    differentiable w.r.t. conv bias, gates weight/bias and Lambda — one kernel each way."""
    return _PhantomH0.apply(conv_bias, gates_w, gates_b, Lambda, pad_len)


class _ScanCL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, h0):
        L.require_cuda(a, b, h0)
        B, T, C = a.shape
        assert b.shape == (B, T, C) and a.dtype == b.dtype
        dt = L.dtype_tag(a)
        a, b = L.as_cl(a), L.as_cl(b)
        h0f, h0_bs = None, 0
        if h0 is not None:
            h0f = h0.detach().float().contiguous()
            h0_bs = C if h0f.dim() == 2 else 0
        h = torch.empty((B, T, C), dtype=a.dtype, device=a.device)
        L.check(L.load().bdlru_scan_cl_fwd(L.view3(a), L.view3(b), L.ptr(h0f), h0_bs, L.view3(h), B, T, C, dt,
                                           L.stream_ptr(a)))
        ctx.save_for_backward(a, h0f, h)
        ctx.h0_bs = h0_bs
        ctx.h0_dtype = h0.dtype if h0 is not None else None
        return h

    @staticmethod
    def backward(ctx, grad):
        a, h0f, h = ctx.saved_tensors
        B, T, C = a.shape
        grad = L.as_cl(grad.to(a.dtype))
        da, db = torch.empty_like(h), torch.empty_like(h)
        dh0 = torch.empty_like(h0f) if h0f is not None else None
        lib = L.load()
        nws = lib.bdlru_gated_scan_bwd_workspace_bytes(B, T, C)
        ws = _workspace(a.device, nws)
        L.check(lib.bdlru_scan_cl_bwd(L.view3(a), L.ptr(h0f), ctx.h0_bs, L.view3(h), L.view3(grad), L.view3(da),
                                      L.view3(db), L.ptr(dh0), L.ptr(ws), nws, B, T, C, L.dtype_tag(a),
                                      L.stream_ptr(a)))
        return da, db, (dh0.to(ctx.h0_dtype) if dh0 is not None else None)


def scan_channel_last(a, b, h0=None):
    """Raw scan h_t = a_t*h_{t-1} + b_t on channel-last [B, T, C] (time along dim 1)."""
    return _ScanCL.apply(a, b, h0)


class _Conv1dCL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, silu):
        L.require_cuda(x, weight, bias)
        B, T, C = x.shape
        assert weight.shape[0] == C and weight.dim() == 2
        W = weight.shape[1]
        x = L.as_cl(x)
        wf = weight.detach().float().contiguous()
        bf = bias.detach().float().contiguous() if bias is not None else None
        y = torch.empty((B, T, C), dtype=x.dtype, device=x.device)
        L.check(L.load().bdlru_conv1d_fwd(L.view3(x), L.ptr(wf), L.ptr(bf), L.view3(y), B, T, C, W, int(silu),
                                          L.dtype_tag(x), L.stream_ptr(x)))
        ctx.save_for_backward(x, wf, bf)
        ctx.silu = bool(silu)
        ctx.w_dtype = weight.dtype
        ctx.b_dtype = bias.dtype if bias is not None else None
        return y

    @staticmethod
    def backward(ctx, grad_y):
        x, wf, bf = ctx.saved_tensors
        B, T, C = x.shape
        W = wf.shape[1]
        grad_y = L.as_cl(grad_y.to(x.dtype))
        dx = torch.empty((B, T, C), dtype=x.dtype, device=x.device)
        dw = torch.empty_like(wf)
        db = torch.empty(C, dtype=torch.float32, device=x.device) if bf is not None else None
        lib = L.load()
        nws = lib.bdlru_conv1d_bwd_workspace_bytes(B, T, C, W)
        ws = _workspace(x.device, nws)
        L.check(lib.bdlru_conv1d_bwd(L.view3(x), L.ptr(wf), L.ptr(bf), L.view3(grad_y), L.view3(dx), L.ptr(dw),
                                     L.ptr(db), L.ptr(ws), nws, B, T, C, W, int(ctx.silu), L.dtype_tag(x),
                                     L.stream_ptr(x)))
        return dx, dw.to(ctx.w_dtype), (db.to(ctx.b_dtype) if db is not None else None), None


def causal_conv1d_channel_last(x, weight, bias=None, silu=True):
    """y_t = act(bias + sum_j weight[:, j] * x_{t-(W-1)+j}) on channel-last [B, T, C]; weight [C, W], W <= 4."""
    return _Conv1dCL.apply(x, weight, bias, silu)


class _BDLRUBlock(torch.autograd.Function):
    """The middle of GatedRecurrentLayer.forward (RecBLR.py:174-206 without the two projections) as ONE autograd node:
        x, z = xz.chunk(2);  x' = silu(conv(x));  (r | i) = gates(x');  y = silu(z) * scan(gate_math(x', r, i), h0)
    Same kernels as the separate ops; what it removes is autograd glue: the gradient of xz is ONE buffer whose halves
    are written in place by the conv backward (dx) and the scan backward (dz) instead of being concatenated, and the two
    contributions to dx' (scan and gates GEMM) are summed inside the GEMM (addmm) instead of by an extra add kernel."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=None)
    def forward(ctx, xz, conv_w, conv_b, gates_w, gates_b, Lambda, h0, use_conv):
        L.require_cuda(xz, gates_w, gates_b, Lambda, h0)
        B, T, C2 = xz.shape
        C = C2 // 2
        lib = L.load()
        xz = L.as_cl(xz)
        dt = L.dtype_tag(xz)
        st = L.stream_ptr(xz)
        x, z = xz[..., :C], xz[..., C:]
        if use_conv:
            wf = conv_w.detach().float().contiguous()
            bf = conv_b.detach().float().contiguous()
            xc = torch.empty((B, T, C), dtype=xz.dtype, device=xz.device)
            L.check(lib.bdlru_conv1d_fwd(L.view3(x), L.ptr(wf), L.ptr(bf), L.view3(xc), B, T, C, wf.shape[1], 1, dt, st))
        else:
            wf = bf = None
            xc = x
        gw = gates_w.detach().to(xz.dtype)
        gb = gates_b.detach().to(xz.dtype)
        with torch.autocast("cuda", enabled=False):
            ri = torch.nn.functional.linear(xc, gw, gb)          # [B, T, 2C], cuBLAS
        Lf = Lambda.detach().float().contiguous()
        h0f = h0.detach().float().contiguous() if h0 is not None else None
        h = torch.empty((B, T, C), dtype=xz.dtype, device=xz.device)
        y = torch.empty_like(h)
        L.check(lib.bdlru_gated_scan_fwd(L.view3(xc), L.view3(ri[..., :C]), L.view3(ri[..., C:]), L.ptr(Lf), L.ptr(h0f), 0,
                                         L.view3(z), L.view3(h), L.view3(y), B, T, C, dt, st))
        ctx.save_for_backward(xz, xc, ri, h, wf, bf, gw, Lf, h0f)
        ctx.meta = (use_conv, conv_w.dtype if conv_w is not None else None, conv_b.dtype if conv_b is not None else None,
                    gates_w.dtype, gates_b.dtype, Lambda.dtype, h0.dtype if h0 is not None else None)
        return y

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        xz, xc, ri, h, wf, bf, gw, Lf, h0f = ctx.saved_tensors
        use_conv, cw_dt, cb_dt, gw_dt, gb_dt, lam_dt, h0_dt = ctx.meta
        B, T, C2 = xz.shape
        C = C2 // 2
        lib = L.load()
        dt = L.dtype_tag(xz)
        st = L.stream_ptr(xz)
        x, z = xz[..., :C], xz[..., C:]
        dy = L.as_cl(dy.to(xz.dtype))
        dxz = torch.empty((B, T, C2), dtype=xz.dtype, device=xz.device)     # (dx | dz), filled in place
        dxc = torch.empty((B, T, C), dtype=xz.dtype, device=xz.device)
        dri = torch.empty((B, T, C2), dtype=xz.dtype, device=xz.device)
        dLambda = torch.empty(C, dtype=torch.float32, device=xz.device)
        dh0 = torch.empty_like(h0f) if h0f is not None else None
        nws = lib.bdlru_gated_scan_bwd_workspace_bytes(B, T, C)
        ws = _workspace(xz.device, nws)
        L.check(lib.bdlru_gated_scan_bwd(L.view3(xc), L.view3(ri[..., :C]), L.view3(ri[..., C:]), L.ptr(Lf), L.ptr(h0f), 0,
                                         L.view3(z), L.view3(h), L.view3(dy), L.view3(dxc), L.view3(dri[..., :C]),
                                         L.view3(dri[..., C:]), L.view3(dxz[..., C:]), L.ptr(dLambda), L.ptr(dh0), L.ptr(ws),
                                         nws, B, T, C, dt, st))
        dri2, xc2 = dri.view(-1, C2), xc.reshape(-1, C)
        # d x' = scan part + dri @ W_gates, summed by the GEMM epilogue
        dxc_tot = torch.addmm(dxc.view(-1, C), dri2, gw).view(B, T, C)
        dgw = _weight_grad(dri2, xc2, gw_dt)
        vw = 4 if dri2.dtype == torch.float32 else 8
        dgb = colsum(dri2) if (C2 % vw == 0 and C2 // vw <= 256) else dri2.float().sum(0)
        if use_conv:
            dw = torch.empty_like(wf)
            db = torch.empty(C, dtype=torch.float32, device=xz.device)
            nwc = lib.bdlru_conv1d_bwd_workspace_bytes(B, T, C, wf.shape[1])
            wsc = _workspace(xz.device, nwc)
            L.check(lib.bdlru_conv1d_bwd(L.view3(x), L.ptr(wf), L.ptr(bf), L.view3(dxc_tot), L.view3(dxz[..., :C]), L.ptr(dw),
                                         L.ptr(db), L.ptr(wsc), nwc, B, T, C, wf.shape[1], 1, dt, st))
            dcw, dcb = dw.to(cw_dt), db.to(cb_dt)
        else:
            dxz[..., :C].copy_(dxc_tot)
            dcw = dcb = None
        return (dxz, dcw, dcb, dgw, dgb.to(gb_dt), dLambda.to(lam_dt),
                dh0.to(h0_dt) if dh0 is not None else None, None)


def bdlru_core_supported(xz, C):
    """True when the fused inference kernel (csrc/fused_core.cu) takes this input: bf16, contiguous [B, T, 2C], C = 128."""
    return (xz.dtype == torch.bfloat16 and xz.is_cuda and xz.is_contiguous() and xz.data_ptr() % 16 == 0
            and bool(L.load().bdlru_core_fwd_supported(int(C), L.BF16)))


@torch.no_grad()
def bdlru_core_fused(xz, conv_w, conv_b, gates_w, gates_b, Lambda, h0=None, use_conv=True):
    """INFERENCE form of bdlru_block as ONE tcgen05 kernel: conv + SiLU, the gates GEMM on the tensor
    cores, gate math, the recurrence and the z-gate, with x' and the gate pre-activations never written to HBM.
    xz bf16 contiguous [B, T, 2C] with C = 128; returns y bf16 [B, T, C].  Not differentiable (no saved activations)."""
    L.require_cuda(xz, gates_w, gates_b, Lambda)
    B, T, C2 = xz.shape
    C = C2 // 2
    assert bdlru_core_supported(xz, C), "bdlru_core_fused: bf16 contiguous [B, T, 256] input required"
    gw = gates_w.detach().to(torch.bfloat16).contiguous()
    gb = gates_b.detach().float().contiguous()
    lam = Lambda.detach().float().contiguous()
    h0f = h0.detach().float().contiguous() if h0 is not None else None
    cw = conv_w.detach().float().contiguous() if use_conv else None
    cb = conv_b.detach().float().contiguous() if use_conv else None
    assert cw is None or cw.shape == (C, 4), "bdlru_core_fused: conv kernel size 4 is built"
    y = torch.empty((B, T, C), dtype=torch.bfloat16, device=xz.device)
    L.check(L.load().bdlru_core_fwd(L.ptr(xz), L.ptr(cw), L.ptr(cb), L.ptr(gw), L.ptr(gb), L.ptr(lam), L.ptr(h0f), L.ptr(y),
                                    B, T, C, L.stream_ptr(xz)))
    return y


def bdlru_layer_supported(x, C, conv_width):
    """True when the full-chain inference kernel (csrc/fused_layer.cu) takes this layer input: bf16 contiguous [B, T, 64]."""
    return (x.dtype == torch.bfloat16 and x.is_cuda and x.is_contiguous() and x.data_ptr() % 16 == 0 and x.dim() == 3
            and bool(L.load().bdlru_layer_fwd_supported(int(x.shape[-1]), int(C), int(conv_width), L.BF16)))


@torch.no_grad()
def bdlru_layer_fused(x, in_w, conv_w, conv_b, gates_w, gates_b, Lambda, h0, out_w, ln_gamma, ln_beta, eps, use_conv=True):
    """INFERENCE form of the first half of RecurrentLayer.forward — in-projection, conv + SiLU, gates GEMM, gate math,
    recurrence, z-gate, out-projection, + residual, LayerNorm (RecBLR.py:140-142, 170-206) — as ONE tcgen05 kernel: reads
    x [B, T, 64] bf16 once, writes the post-LayerNorm [B, T, 64] bf16; nothing in between touches HBM."""
    L.require_cuda(x, in_w, gates_w, out_w)
    B, T, D = x.shape
    C = gates_w.shape[1]
    bf = lambda w: w.detach().to(torch.bfloat16).contiguous()
    fl = lambda w: w.detach().float().contiguous()
    iw, gw, ow = bf(in_w), bf(gates_w), bf(out_w)
    assert iw.shape == (2 * C, D) and gw.shape == (2 * C, C) and ow.shape == (D, C)
    cw = fl(conv_w) if use_conv else None
    cb = fl(conv_b) if use_conv else None
    h0f = fl(h0) if h0 is not None else None
    out = torch.empty((B, T, D), dtype=torch.bfloat16, device=x.device)
    L.check(L.load().bdlru_layer_fwd(L.ptr(x), L.ptr(iw), L.ptr(cw), L.ptr(cb), L.ptr(gw), L.ptr(fl(gates_b)), L.ptr(fl(Lambda)),
                                     L.ptr(h0f), L.ptr(ow), L.ptr(fl(ln_gamma)), L.ptr(fl(ln_beta)), float(eps), L.ptr(out),
                                     B, T, D, C, L.stream_ptr(x)))
    return out


def bdlru_block(xz, conv_w, conv_b, gates_w, gates_b, Lambda, h0=None, use_conv=True):
    """y = silu(z) * BD-LRU(silu(conv(x))) for xz = (x | z) [B, T, 2C] — conv, gates GEMM, gate math, scan and z-gate of
    RecBLR.py:174-206 as one autograd node (see _BDLRUBlock).  conv_w [C, W], gates_w [2C, C]; autocast aware."""
    return _BDLRUBlock.apply(xz, conv_w, conv_b, gates_w, gates_b, Lambda, h0, use_conv)


class _EmbedLN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, table, gamma, beta, eps, p, seed, padding_idx, seed_dev, out_dtype):
        L.require_cuda(ids, table, gamma, beta)
        assert ids.dtype == torch.int64 and table.dim() == 2
        n_items, D = table.shape
        ids_c = ids.contiguous()
        tab = table.detach().contiguous()
        gf, bf = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        n = ids_c.numel()
        out_dtype = out_dtype or table.dtype
        out = torch.empty((*ids.shape, D), dtype=out_dtype, device=table.device)
        mean = torch.empty(n, dtype=torch.float32, device=table.device)
        rstd = torch.empty_like(mean)
        L.check(L.load().bdlru_embed_ln_fwd(L.ptr(ids_c), L.ptr(tab), L.ptr(gf), L.ptr(bf), L.ptr(out), L.ptr(mean),
                                            L.ptr(rstd), n, n_items, D, float(eps), float(p), int(seed),
                                            L.ptr(seed_dev), L.dtype_tag(tab), L.dtype_tag(out), L.stream_ptr(tab)))
        ctx.save_for_backward(ids_c, tab, gf, mean, rstd)
        ctx.seed_dev = seed_dev
        ctx.args = (float(p), int(seed), int(padding_idx), gamma.dtype, beta.dtype, out_dtype)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        ids_c, tab, gf, mean, rstd = ctx.saved_tensors
        p, seed, padding_idx, g_dtype, b_dtype, out_dtype = ctx.args
        n_items, D = tab.shape
        n = ids_c.numel()
        grad_out = grad_out.to(out_dtype).contiguous()
        dtable = torch.zeros((n_items, D), dtype=torch.float32, device=tab.device)
        dgamma = torch.empty(D, dtype=torch.float32, device=tab.device)
        dbeta = torch.empty_like(dgamma)
        lib = L.load()
        nws = lib.bdlru_embed_ln_bwd_workspace_bytes(n, D)
        ws = _workspace(tab.device, nws)
        L.check(lib.bdlru_embed_ln_bwd(L.ptr(ids_c), L.ptr(tab), L.ptr(gf), L.ptr(grad_out), L.ptr(mean), L.ptr(rstd),
                                       L.ptr(dtable), L.ptr(dgamma), L.ptr(dbeta), L.ptr(ws), nws, n, n_items, D, p,
                                       seed, L.ptr(ctx.seed_dev), padding_idx, L.dtype_tag(tab), L.dtype_tag(grad_out),
                                       L.stream_ptr(tab)))
        return None, dtable.to(tab.dtype), dgamma.to(g_dtype), dbeta.to(b_dtype), None, None, None, None, None, None


def embed_layernorm(ids, table, gamma, beta, eps=1e-12, dropout_p=0.0, seed=0, padding_idx=-1, seed_dev=None,
                    out_dtype=None):
    """LayerNorm(dropout(table[ids])) in one kernel (RecBLR.py:76-78).  ids int64 [...]; returns [..., D] in
    table.dtype (or out_dtype = bf16 for an fp32 table).  Rows equal to padding_idx receive no gather gradient (nn.Embedding(padding_idx=0) semantics).
    seed_dev: optional int64[1] CUDA tensor added to `seed` on the device (the mask must NOT change between this
    call's forward and backward: bump it once per step, before the forward)."""
    return _EmbedLN.apply(ids, table, gamma, beta, eps, dropout_p, seed, padding_idx, seed_dev, out_dtype)


def scatter_add_rows(ids, rows, dst, row_lo, row_hi, padding_idx=-1):
    """dst[(id - row_lo)] += rows[n] for every token n with row_lo <= ids[n] < row_hi and ids[n] != padding_idx
    (dst fp32 [row_hi - row_lo or more, D]; rows fp32/bf16 [n, D]) — the owner-side half of the sharded embedding
    gradient."""
    L.require_cuda(ids, rows, dst)
    assert ids.dtype == torch.int64 and rows.dim() == 2 and dst.dtype == torch.float32
    ids_c, rows_c = ids.reshape(-1).contiguous(), rows.contiguous()
    assert rows_c.shape[0] == ids_c.numel() and dst.is_contiguous() and dst.shape[1] == rows_c.shape[1]
    assert dst.shape[0] >= row_hi - row_lo
    L.check(L.load().bdlru_scatter_add_rows(L.ptr(ids_c), L.ptr(rows_c), ids_c.numel(), rows_c.shape[1],
                                            L.dtype_tag(rows_c), int(row_lo), int(row_hi), int(padding_idx), L.ptr(dst),
                                            L.stream_ptr(dst)))
    return dst


def colsum(x2d):
    """fp32 column sums of a row-major [rows, cols] fp32/bf16 matrix (one streaming pass + deterministic reduce)."""
    L.require_cuda(x2d)
    assert x2d.dim() == 2 and x2d.stride(1) == 1
    rows, cols = x2d.shape
    out = torch.empty(cols, dtype=torch.float32, device=x2d.device)
    lib = L.load()
    nws = lib.bdlru_colsum_workspace_bytes(rows, cols)
    ws = _workspace(x2d.device, nws)
    L.check(lib.bdlru_colsum(L.ptr(x2d), rows, cols, x2d.stride(0), L.dtype_tag(x2d), L.ptr(out), L.ptr(ws), nws,
                             L.stream_ptr(x2d)))
    return out


def _weight_grad(dy2, x2, want_dtype):
    """dW = dy2^T x2 in `want_dtype`; bf16 operands with an fp32 master weight accumulate and WRITE in fp32 (no bf16 rounding
    of the gradient, no cast kernel)."""
    with torch.autocast("cuda", enabled=False):
        if dy2.dtype == torch.bfloat16 and want_dtype == torch.float32:
            return torch.mm(dy2.t(), x2, out_dtype=torch.float32)
        return (dy2.t() @ x2).to(want_dtype)


class _LinearBias(torch.autograd.Function):
    """y = x W^T + b with the reference's nn.Linear semantics; GEMMs stay cuBLAS, only the bias gradient (a column sum
    ATen computes with a slow generic reduction) goes through bdlru_colsum."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=None)
    def forward(ctx, x, weight, bias):
        ctx.w_dtype = weight.dtype
        if torch.is_autocast_enabled("cuda"):
            dt = torch.get_autocast_dtype("cuda")
            x, weight, bias_c = x.to(dt), weight.to(dt), bias.to(dt)
        else:
            bias_c = bias
        ctx.save_for_backward(x, weight)
        ctx.bias_dtype = bias.dtype
        with torch.autocast("cuda", enabled=False):
            return torch.nn.functional.linear(x, weight, bias_c)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        x2 = x.reshape(-1, x.shape[-1])
        dx = (dy2 @ weight).view_as(x)
        dw = _weight_grad(dy2, x2, ctx.w_dtype)
        vw = 4 if dy2.dtype == torch.float32 else 8
        if dy2.shape[1] % vw == 0 and dy2.shape[1] // vw <= 256 and dy2.dtype in (torch.float32, torch.bfloat16):
            db = colsum(dy2)
        else:
            db = dy2.float().sum(0)
        return dx, dw, db.to(ctx.bias_dtype)


class _LinearTap(torch.autograd.Function):
    """(x W^T [+ b], x): a Linear whose INPUT is also handed on as a second output (the residual branch of
    RecurrentLayer.forward RecBLR.py:141-142 and FeedForward.forward RecBLR.py:219-225).  With x consumed by this single
    node, the two contributions to dL/dx (through the GEMM and through the residual) are summed inside the GEMM
    (addmm) instead of by autograd's gradient-accumulation add kernel."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=None)
    def forward(ctx, x, weight, bias):
        ctx.w_dtype = weight.dtype
        ctx.b_dtype = bias.dtype if bias is not None else None
        xc, w, b = x, weight, bias
        if torch.is_autocast_enabled("cuda"):
            dt = torch.get_autocast_dtype("cuda")
            xc, w = x.to(dt), weight.to(dt)
            b = bias.to(dt) if bias is not None else None
        ctx.save_for_backward(xc, w)
        ctx.x_dtype = x.dtype
        with torch.autocast("cuda", enabled=False):
            y = torch.nn.functional.linear(xc, w, b)
        return y, x

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dy, dtap):
        xc, w = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        x2 = xc.reshape(-1, xc.shape[-1])
        with torch.autocast("cuda", enabled=False):
            if dtap is not None:
                dx = torch.addmm(dtap.to(dy2.dtype).reshape(-1, xc.shape[-1]), dy2, w).view_as(xc)
            else:
                dx = (dy2 @ w).view_as(xc)
        dw = _weight_grad(dy2, x2, ctx.w_dtype)
        db = None
        if ctx.b_dtype is not None:
            vw = 4 if dy2.dtype == torch.float32 else 8
            ok = dy2.shape[1] % vw == 0 and dy2.shape[1] // vw <= 256 and dy2.dtype in (torch.float32, torch.bfloat16)
            db = (colsum(dy2) if ok else dy2.float().sum(0)).to(ctx.b_dtype)
        return dx.to(ctx.x_dtype), dw, db


def linear_tap(x, weight, bias=None):
    """Returns (linear(x), x) as one autograd node (see _LinearTap); autocast aware."""
    L.require_cuda(x, weight)
    return _LinearTap.apply(x, weight, bias)


def linear_bias(x, weight, bias):
    """Drop-in for nn.Linear(...)(x) when the layer has a bias (autocast aware)."""
    L.require_cuda(x, weight, bias)
    return _LinearBias.apply(x, weight, bias)


class _SiluDropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, seed, seed_dev):
        L.require_cuda(x)
        xc = x.contiguous()
        out = torch.empty_like(xc)
        L.check(L.load().bdlru_silu_dropout_fwd(L.ptr(xc), L.ptr(out), xc.numel(), float(p), int(seed), L.ptr(seed_dev),
                                                L.dtype_tag(xc), L.stream_ptr(xc)))
        ctx.save_for_backward(xc)
        ctx.args = (float(p), int(seed), seed_dev)
        return out

    @staticmethod
    def backward(ctx, dy):
        (xc,) = ctx.saved_tensors
        p, seed, seed_dev = ctx.args
        dy = dy.to(xc.dtype).contiguous()
        dx = torch.empty_like(xc)
        L.check(L.load().bdlru_silu_dropout_bwd(L.ptr(xc), L.ptr(dy), L.ptr(dx), xc.numel(), p, seed, L.ptr(seed_dev),
                                                L.dtype_tag(xc), L.stream_ptr(xc)))
        return dx, None, None, None


def silu_dropout(x, dropout_p=0.0, seed=0, seed_dev=None):
    """dropout(silu(x)) in one kernel (RecBLR.py:219-221); numel % 8 == 0."""
    return _SiluDropout.apply(x, dropout_p, seed, seed_dev)


class _AddLN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, res, gamma, beta, eps, p, seed, seed_dev):
        L.require_cuda(x, res, gamma, beta)
        assert x.shape == res.shape
        D = x.shape[-1]
        dt = x.dtype
        xc, rc = x.contiguous(), res.to(dt).contiguous()
        gf, bf = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        n = xc.numel() // D
        out = torch.empty_like(xc)
        mean = torch.empty(n, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        L.check(L.load().bdlru_add_ln_fwd(L.ptr(xc), L.ptr(rc), L.ptr(gf), L.ptr(bf), L.ptr(out), L.ptr(mean), L.ptr(rstd),
                                          n, D, float(eps), float(p), int(seed), L.ptr(seed_dev), L.dtype_tag(xc),
                                          L.stream_ptr(xc)))
        ctx.save_for_backward(xc, rc, gf, mean, rstd)
        ctx.seed_dev = seed_dev
        ctx.args = (float(p), int(seed), res.dtype, gamma.dtype, beta.dtype)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        xc, rc, gf, mean, rstd = ctx.saved_tensors
        p, seed, res_dtype, g_dtype, b_dtype = ctx.args
        D = xc.shape[-1]
        n = xc.numel() // D
        grad_out = grad_out.to(xc.dtype).contiguous()
        dres = torch.empty_like(xc)
        dx = torch.empty_like(xc) if p > 0.0 else dres
        dgamma = torch.empty(D, dtype=torch.float32, device=xc.device)
        dbeta = torch.empty_like(dgamma)
        lib = L.load()
        nws = lib.bdlru_add_ln_bwd_workspace_bytes(n, D)
        ws = _workspace(xc.device, nws)
        L.check(lib.bdlru_add_ln_bwd(L.ptr(xc), L.ptr(rc), L.ptr(gf), L.ptr(grad_out), L.ptr(mean), L.ptr(rstd), L.ptr(dx),
                                     L.ptr(dres), L.ptr(dgamma), L.ptr(dbeta), L.ptr(ws), nws, n, D, p, seed,
                                     L.ptr(ctx.seed_dev), L.dtype_tag(xc), L.stream_ptr(xc)))
        return dx, dres.to(res_dtype), dgamma.to(g_dtype), dbeta.to(b_dtype), None, None, None, None


def add_dropout_layernorm(x, residual, gamma, beta, eps=1e-12, dropout_p=0.0, seed=0, seed_dev=None):
    """LayerNorm(dropout(x) + residual) in one kernel (RecBLR.py:142, 221-225); output in x.dtype (the residual is
    read in that dtype too), fp32 statistics.  D % 4 == 0, D <= 512."""
    return _AddLN.apply(x, residual, gamma, beta, eps, dropout_p, seed, seed_dev)


# ----------------------------------------------------------------------------- full-sort scoring / CE (tcgen05)
def fullsort_supported(D):
    """Shapes the tcgen05 full-sort kernels take: bf16 operands, D a multiple of 64 up to 256."""
    return D % 64 == 0 and D <= 256 and bool(L.load().bdlru_fullsort_available())


def _bf16_rows(t):
    """bf16, row-major, 16-byte aligned copy/view of a [rows, D] matrix (the operand format of the tensor cores)."""
    t = t.detach()
    if t.dtype != torch.bfloat16:
        t = t.to(torch.bfloat16)
    return t.contiguous()


BIAS_COLS = 64  # K granule of the full-sort kernels: an item bias rides the tensor cores as one extra granule


def fullsort_bias_supported(D):
    return fullsort_supported(D + BIAS_COLS)


def _augment_with_bias(qb, eb, item_bias):
    """Operands of `q @ table^T + item_bias` (the BERT4Rec scorer, bert4rec.py:230-242) as ONE bf16 GEMM with fp32
    accumulation: q' = [q | 1 1 1 0...], table' = [table | b_hi b_mid b_lo 0...] with b = b_hi + b_mid + b_lo the EXACT
    three-way bf16 split of the fp32 bias (3 x 8 mantissa bits), so the epilogues (top-k, softmax statistics, softmax
    backward) need no bias path and d(bias) falls out as column D of d(table')."""
    assert item_bias.dim() == 1 and item_bias.shape[0] == eb.shape[0]
    b = item_bias.detach().float()
    hi = b.to(torch.bfloat16)
    r1 = b - hi.float()
    mid = r1.to(torch.bfloat16)
    lo = (r1 - mid.float()).to(torch.bfloat16)
    B, D = qb.shape
    qa = qb.new_zeros((B, D + BIAS_COLS))
    qa[:, :D] = qb
    qa[:, D:D + 3] = 1
    ea = eb.new_zeros((eb.shape[0], D + BIAS_COLS))
    ea[:, :D] = eb
    ea[:, D], ea[:, D + 1], ea[:, D + 2] = hi, mid, lo
    return qa, ea


def fullsort_topk(q, table, k, mask_id=0, id_offset=0, item_bias=None, out=None):
    """Fused full-sort scoring + top-k (RecBLR.py:114-122 + RecBole's `scores[:, 0] = -inf; torch.topk`): returns
    (scores fp32 [B, k], ids int32 [B, k]) of q @ table^T (+ item_bias [n_rows], bert4rec.py:230-242) without forming
    [B, n_rows].  Operands are rounded to bf16, accumulated in fp32 on the tcgen05 tensor cores; ordering is score
    descending, LOWEST id first among ties; the row with global id `mask_id` is excluded (-1: none).  Row j of `table`
    has global id id_offset + j (row shards)."""
    L.require_cuda(q, table)
    assert q.dim() == 2 and table.dim() == 2 and q.shape[1] == table.shape[1]
    qb, eb = _bf16_rows(q), _bf16_rows(table)
    if item_bias is not None:
        qb, eb = _augment_with_bias(qb, eb, item_bias)
    B, D = qb.shape
    N = table.shape[0]
    lib = L.load()
    if out is not None:   # caller-owned contiguous [B, k] fp32 / int32 buffers (e.g. halves of one exchange buffer)
        out_s, out_i = out
        assert out_s.shape == (B, k) and out_s.dtype == torch.float32 and out_s.is_contiguous()
        assert out_i.shape == (B, k) and out_i.dtype == torch.int32 and out_i.is_contiguous()
    else:
        out_s = torch.empty((B, k), dtype=torch.float32, device=q.device)
        out_i = torch.empty((B, k), dtype=torch.int32, device=q.device)
    nws = lib.bdlru_fullsort_topk_workspace_bytes(B, N, D, k)
    ws = _workspace(q.device, nws)
    L.check(lib.bdlru_fullsort_topk(L.ptr(qb), L.ptr(eb), B, N, D, k, id_offset, mask_id, L.ptr(out_s), L.ptr(out_i),
                                    L.ptr(ws), nws, L.stream_ptr(q)))
    return out_s, out_i


def topk_merge(cand_scores, cand_ids, k):
    """Keeps the k best of [B, n_lists * k] candidates by (score desc, id asc) — the local step of the NCCL merge."""
    L.require_cuda(cand_scores, cand_ids)
    B, nc = cand_scores.shape
    assert nc % k == 0 and cand_ids.shape == (B, nc) and cand_ids.dtype == torch.int32
    cs, ci = cand_scores.float().contiguous(), cand_ids.contiguous()
    out_s = torch.empty((B, k), dtype=torch.float32, device=cs.device)
    out_i = torch.empty((B, k), dtype=torch.int32, device=cs.device)
    L.check(L.load().bdlru_topk_merge(L.ptr(cs), L.ptr(ci), B, nc // k, k, L.ptr(out_s), L.ptr(out_i), L.stream_ptr(cs)))
    return out_s, out_i


def _check_pos(pos, n_users):
    """The kernels read the positives as `const int64_t*`: anything else would be read out of bounds."""
    assert pos.dim() == 1 and pos.shape[0] == n_users, f"pos must be [{n_users}], got {tuple(pos.shape)}"
    if pos.dtype != torch.int64:
        assert not pos.dtype.is_floating_point, f"pos must hold integer item ids, got {pos.dtype}"
        pos = pos.long()
    return pos.contiguous()


def topk_merge_gathered(packed, k):
    """Merge of the buffer `all_gather_into_tensor` leaves behind when every rank contributes one packed
    [2, B, k] tensor (row 0: fp32 scores, row 1: the int32 ids' bits): packed fp32 [world, 2, B, k], read in place."""
    L.require_cuda(packed)
    world, two, B, kk = packed.shape
    assert two == 2 and kk == k and packed.dtype == torch.float32 and packed.is_contiguous()
    out_s = torch.empty((B, k), dtype=torch.float32, device=packed.device)
    out_i = torch.empty((B, k), dtype=torch.int32, device=packed.device)
    ids_ptr = ctypes.c_void_p(packed.data_ptr() + B * k * 4)
    L.check(L.load().bdlru_topk_merge_strided(L.ptr(packed), ids_ptr, B, world, k, 2 * B * k, k, L.ptr(out_s), L.ptr(out_i),
                                              L.stream_ptr(packed)))
    return out_s, out_i


def fullsort_ce_stats(q, table, pos, id_offset=0, item_bias=None):
    """Per-user (row_max, row_sumexp, pos_logit) of the logits q @ table^T (+ item_bias) over this table (shard), never
    materialised.  pos_logit is written only for users whose positive row lives in this shard (others keep 0)."""
    L.require_cuda(q, table, pos)
    pos = _check_pos(pos, q.shape[0])
    qb, eb = _bf16_rows(q), _bf16_rows(table)
    if item_bias is not None:
        qb, eb = _augment_with_bias(qb, eb, item_bias)
    B, D = qb.shape
    N = table.shape[0]
    lib = L.load()
    row_max = torch.empty(B, dtype=torch.float32, device=q.device)
    row_sum = torch.empty_like(row_max)
    pos_logit = torch.zeros_like(row_max)
    nws = lib.bdlru_fullsort_ce_workspace_bytes(B, N, D)
    ws = _workspace(q.device, nws)
    L.check(lib.bdlru_fullsort_ce_fwd(L.ptr(qb), L.ptr(eb), L.ptr(pos.contiguous()), B, N, D, id_offset, L.ptr(row_max),
                                      L.ptr(row_sum), L.ptr(pos_logit), L.ptr(ws), nws, L.stream_ptr(q)))
    return row_max, row_sum, pos_logit


def fullsort_ce_grads(qb, eb, pos, lse, scale, id_offset=0, scale_dev=None, out_dq=None, out_de=None, want_dq=True,
                      want_de=True):
    """(dQ [B, D], dE [rows, D]) fp32 of scale * sum_b(lse_b - logit_{b,pos_b}) for bf16 qb, eb and the GLOBAL lse;
    the logits are recomputed tile by tile on the tensor cores and never stored.  `scale_dev` (fp32 CUDA scalar) is
    multiplied in on the device (the upstream gradient of the loss: no sync, no extra pass over dE); `out_dq` / `out_de`
    let the caller have the kernel write straight into its own fp32 buffers (e.g. the optimizer's gradient of the table)."""
    L.require_cuda(qb, eb, pos, lse)
    pos = _check_pos(pos, qb.shape[0])
    B, D = qb.shape
    N = eb.shape[0]
    lib = L.load()
    dQ = dE = None
    if want_dq:
        dQ = out_dq if out_dq is not None else torch.empty((B, D), dtype=torch.float32, device=qb.device)
        assert dQ.shape == (B, D) and dQ.dtype == torch.float32 and dQ.is_contiguous()
    if want_de:
        dE = out_de if out_de is not None else torch.empty((N, D), dtype=torch.float32, device=qb.device)
        assert dE.shape == (N, D) and dE.dtype == torch.float32 and dE.is_contiguous()
    if scale_dev is not None:
        scale_dev = scale_dev.detach().reshape(()).float().contiguous()
    nws = lib.bdlru_fullsort_ce_workspace_bytes(B, N, D)
    ws = _workspace(qb.device, nws)
    L.check(lib.bdlru_fullsort_ce_bwd(L.ptr(qb), L.ptr(eb), L.ptr(pos), L.ptr(lse.float().contiguous()), float(scale),
                                      L.ptr(scale_dev), B, N, D, id_offset, L.ptr(dQ), L.ptr(dE), L.ptr(ws), nws,
                                      L.stream_ptr(qb)))
    return dQ, dE


# Softmax reference of the fused CE forward: row maximum over every CE_REFERENCE_STRIDE-th 96-item tile (1 = exact row
# maximum over all items, a full tensor-core pass).  Any reference within ~80 of the true maximum gives the same loss and
# gradients (shift invariance; bf16 / fp32 keep their relative precision), and a sampled maximum over >= 1/16 of a catalog
# is within a few units of the true one; if it ever were not, exp would overflow to inf and the device-side assert below
# fires — the result can be slow (exact pass) or loud, never silently wrong.
CE_REFERENCE_STRIDE = 16


def fullsort_rowmax(qb, eb, tile_stride=1):
    """Row maxima [B] fp32 of qb @ eb^T on the tensor cores, no exponentials (tile_stride > 1: sampled, see above)."""
    L.require_cuda(qb, eb)
    B, D = qb.shape
    N = eb.shape[0]
    lib = L.load()
    out = torch.empty(B, dtype=torch.float32, device=qb.device)
    nws = lib.bdlru_fullsort_rowmax_workspace_bytes(B, N, D)
    ws = _workspace(qb.device, nws)
    L.check(lib.bdlru_fullsort_rowmax(L.ptr(qb), L.ptr(eb), B, N, D, int(max(1, tile_stride)), L.ptr(out), L.ptr(ws), nws,
                                      L.stream_ptr(qb)))
    return out


def fullsort_ce_fwd_dq(qb, eb, ref):
    """Fused forward of the training CE for bf16 qb [B, D], eb [rows, D] and a per-user reference ref [B] fp32:
    (acc [B, D], s [B]) with s_b = sum_j exp(l_bj - ref_b), acc_b = sum_j exp(l_bj - ref_b) eb_j  — one exponential pass
    yields both the softmax denominator and the unnormalised dQ (include/bdlru.h)."""
    L.require_cuda(qb, eb, ref)
    B, D = qb.shape
    N = eb.shape[0]
    lib = L.load()
    acc = torch.empty((B, D), dtype=torch.float32, device=qb.device)
    s = torch.empty(B, dtype=torch.float32, device=qb.device)
    nws = lib.bdlru_fullsort_ce_fwd_dq_workspace_bytes(B, N, D)
    ws = _workspace(qb.device, nws)
    L.check(lib.bdlru_fullsort_ce_fwd_dq(L.ptr(qb), L.ptr(eb), L.ptr(ref.float().contiguous()), B, N, D, L.ptr(acc),
                                         L.ptr(s), L.ptr(ws), nws, L.stream_ptr(qb)))
    return acc, s


def pos_logits(qb, eb, pos, id_offset=0):
    """q_b . E[pos_b] for the positives whose row lives in this (shard of the) table, 0 elsewhere; also the gathered rows
    (zeros for foreign positives).  bf16 x bf16 products are exact in fp32; only the summation order differs from the GEMM."""
    loc = pos - id_offset
    own = (loc >= 0) & (loc < eb.shape[0])
    rows = eb[loc.clamp(0, eb.shape[0] - 1)].float() * own[:, None]
    return (qb.float() * rows).sum(1), rows


class _FullsortCE(torch.autograd.Function):
    """mean_b(logsumexp_n(q_b . E_n) - q_b . E_pos_b) over ALL rows of E (RecBLR.py:99-103), logits never stored.
    When q needs a gradient the forward is the FUSED pass (reference maximum -> one exponential pass giving the softmax
    denominator and the unnormalised dQ), and the backward only runs the dE pass; otherwise the statistics-only kernel."""

    @staticmethod
    def forward(ctx, q, table, pos, item_bias):
        qb, eb = _bf16_rows(q), _bf16_rows(table)
        if item_bias is not None:
            qb, eb = _augment_with_bias(qb, eb, item_bias)
        ctx.meta = (q.dtype, table.dtype, q.shape[1], None if item_bias is None else item_bias.dtype)
        ctx.fused = bool(ctx.needs_input_grad[0])
        if ctx.fused:
            ref = fullsort_rowmax(qb, eb, CE_REFERENCE_STRIDE)
            acc, s = fullsort_ce_fwd_dq(qb, eb, ref)
            torch._assert_async(torch.isfinite(s).all(), "fused CE: the sampled softmax reference is > 80 below a row "
                                "maximum (exp overflow): set ops.CE_REFERENCE_STRIDE = 1")
            lse = ref + torch.log(s)
            pl, prow = pos_logits(qb, eb, pos)
            ctx.save_for_backward(qb, eb, pos, lse, acc, s, prow)
        else:
            m, s, pl = fullsort_ce_stats(qb, eb, pos)
            lse = m + torch.log(s)
            ctx.save_for_backward(qb, eb, pos, lse)
        return (lse - pl).mean()

    @staticmethod
    def backward(ctx, grad_loss):
        qd, ed, D, bd = ctx.meta
        # dloss/dlogit = (softmax - onehot) / B, times the upstream scalar (read by the kernels on the device: no sync and
        # no extra pass over the [n_items, D] gradient)
        if ctx.fused:
            qb, eb, pos, lse, acc, s, prow = ctx.saved_tensors
            dQ = (acc / s[:, None] - prow) * (grad_loss.float() / qb.shape[0])
            dE = None
            if ctx.needs_input_grad[1] or ctx.needs_input_grad[3]:
                _, dE = fullsort_ce_grads(qb, eb, pos, lse, 1.0 / qb.shape[0], scale_dev=grad_loss, want_dq=False)
        else:
            qb, eb, pos, lse = ctx.saved_tensors
            dQ, dE = fullsort_ce_grads(qb, eb, pos, lse, 1.0 / qb.shape[0], scale_dev=grad_loss)
        if bd is None:
            return dQ.to(qd), (dE.to(ed) if dE is not None else None), None, None
        # augmented operands: the bias gradient is the column of d(table') that multiplies q' = 1
        return dQ[:, :D].to(qd), dE[:, :D].to(ed), None, dE[:, D].to(bd)


def fullsort_cross_entropy(q, table, pos, item_bias=None):
    """Fused full-softmax cross-entropy over every row of `table` (incl. the pad row 0), mean over the
    batch: bf16 operands, fp32 accumulation and statistics, [B, n_items] never materialised in forward or backward.
    `item_bias` [n_rows] adds BERT4Rec's `output_bias` to the logits (bert4rec.py:200-213; select the masked positions
    before the call — the reference's `sum(loss * targets) / sum(targets)` is the mean over the rows with target 1)."""
    L.require_cuda(q, table, pos)
    return _FullsortCE.apply(q, table, _check_pos(pos, q.shape[0]), item_bias)
