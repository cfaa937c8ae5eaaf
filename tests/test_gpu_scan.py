"""GPU parity: CUDA scan / gated-scan / conv kernels (through the C ABI) vs the float64 oracle.
Tolerance (north_star): fp32 accumulate, max relative error <= 1e-4 against a float64 sequential scan."""
import numpy as np
import pytest
import torch

from oracle import bdlru_oracle as O
from tests.util import cuda, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4       # fp32 I/O
TOL_BF16 = 2e-2  # bf16 I/O: inputs are rounded identically for the oracle, outputs carry one bf16 rounding


def _ops():
    from datamining_recblr_b200 import ops
    return ops


@pytest.mark.parametrize("B,C,T", [(2, 3, 1), (2, 5, 7), (3, 4, 50), (2, 8, 128), (2, 4, 200), (1, 2, 1000),
                                   (4, 16, 256), (2, 3, 4096), (1, 1, 33), (300, 7, 36)])
@pytest.mark.parametrize("regime", ["mid", "near_one", "near_zero"])
def test_parallel_scan_bct(B, C, T, regime):
    rng = np.random.default_rng(B * 1000 + C * 10 + T)
    lo, hi = {"mid": (0.5, 1.0), "near_one": (0.999, 1.0), "near_zero": (0.0, 0.05)}[regime]
    a = rng.uniform(lo, hi, (B, C, T)).astype(np.float32)
    b = rng.normal(size=(B, C, T)).astype(np.float32)
    g = rng.normal(size=(B, C, T)).astype(np.float32)
    h_ref = O.scan_fwd(a.astype(np.float64), b.astype(np.float64))
    da_ref, db_ref, _ = O.scan_bwd(a.astype(np.float64), h_ref, g.astype(np.float64))
    ta, tb = cuda(a, requires_grad=True), cuda(b, requires_grad=True)
    h = _ops().parallel_scan(ta, tb)
    h.backward(cuda(g))
    assert rel_err(h, h_ref) <= TOL
    assert rel_err(tb.grad, db_ref) <= TOL
    assert rel_err(ta.grad, da_ref) <= TOL


def test_parallel_scan_contract_errors():
    ops = _ops()
    a = torch.rand(2, 3, 8, device="cuda")
    with pytest.raises(AssertionError):   # parallel_scan.py:88-89: contiguity asserts
        ops.parallel_scan(a.transpose(1, 2), a.transpose(1, 2))
    with pytest.raises(AssertionError):   # parallel_scan.py:87: shape assert
        ops.parallel_scan(a, a[:, :, :4].contiguous())
    with pytest.raises(Exception):        # no CPU fallback
        ops.parallel_scan(a.cpu(), a.cpu())


def _gated_inputs(rng, B, T, C, dtype):
    xp = rng.normal(size=(B, T, C))
    ri = rng.normal(size=(B, T, 2 * C)) * 1.5
    lam = O.lambda_init(C) + rng.normal(size=C) * 0.3
    h0 = rng.normal(size=C)
    h0b = rng.normal(size=(B, C))
    z = rng.normal(size=(B, T, C))
    g = rng.normal(size=(B, T, C))
    if dtype == torch.bfloat16:  # round inputs the way the kernel will see them
        rd = lambda v: torch.tensor(v, dtype=torch.float32).to(torch.bfloat16).double().numpy()
        xp, ri, z, g = rd(xp), rd(ri), rd(z), rd(g)
    return xp, ri, lam.astype(np.float32).astype(np.float64), h0.astype(np.float32).astype(np.float64), \
        h0b.astype(np.float32).astype(np.float64), z, g


@pytest.mark.parametrize("B,T,C", [(2, 1, 4), (3, 5, 8), (2, 50, 128), (2, 200, 128), (1, 37, 64), (2, 130, 256),
                                   (3, 64, 12), (5, 33, 192), (160, 50, 128), (2, 1000, 64),
                                   # large B*C: the sequential bf16 variant (one thread per (row, channel vector)), V = 4 / 2
                                   (4096, 9, 128), (2048, 6, 128), (4100, 1, 128), (1024, 11, 256)])
@pytest.mark.parametrize("h0_mode", ["none", "bcast", "batch"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gated_scan(B, T, C, h0_mode, dtype):
    rng = np.random.default_rng(B * 100000 + T * 100 + C)
    xp, ri, lam, h0, h0b, z, g = _gated_inputs(rng, B, T, C, dtype)
    h0_np = {"none": None, "bcast": h0, "batch": h0b}[h0_mode]
    r_np, i_np = ri[..., :C], ri[..., C:]
    h_ref = O.gated_scan_fwd(xp, r_np, i_np, lam, h0_np)
    if h0_mode == "batch":
        # oracle's dh0 sums over batch; get the per-batch grad from the per-batch call
        dxp, dr, di, dlam, _ = O.gated_scan_bwd(xp, r_np, i_np, lam, h0_np, g)
    else:
        dxp, dr, di, dlam, dh0 = O.gated_scan_bwd(xp, r_np, i_np, lam, h0_np, g)
    tol = TOL if dtype == torch.float32 else TOL_BF16
    txp = cuda(xp, dtype, True)
    tri = cuda(ri, dtype, True)          # r, i are strided views of one [B,T,2C] tensor (RecBLR.py:196)
    tlam = cuda(lam, torch.float32, True)
    th0 = cuda(h0_np, torch.float32, True) if h0_np is not None else None
    r_t, i_t = tri.chunk(2, dim=-1)
    h = _ops().gated_scan(txp, r_t, i_t, tlam, th0)
    assert h.dtype == dtype
    h.backward(cuda(g, dtype))
    assert rel_err(h, h_ref) <= tol
    assert rel_err(txp.grad, dxp) <= tol
    assert rel_err(tri.grad[..., :C], dr) <= tol
    assert rel_err(tri.grad[..., C:], di) <= tol
    assert rel_err(tlam.grad, dlam) <= (5e-4 if dtype == torch.float32 else 5e-2)
    if h0_mode == "bcast":
        assert rel_err(th0.grad, dh0) <= tol
    if h0_mode == "batch":
        a, _ = O.gate_math(xp, r_np, i_np, lam)
        _, dtok, _ = O.scan_bwd(np.swapaxes(a, 1, 2), np.swapaxes(h_ref, 1, 2), np.swapaxes(g, 1, 2))
        assert rel_err(th0.grad, a[:, 0, :] * dtok[:, :, 0]) <= tol


@pytest.mark.parametrize("B,T,C", [(2, 5, 8), (2, 50, 128), (3, 200, 128), (2, 67, 64), (4096, 9, 128), (2048, 6, 128),
                                   (8192, 3, 64), (1024, 5, 256)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gated_scan_fused_zgate(B, T, C, dtype):
    rng = np.random.default_rng(7 + T)
    xp, ri, lam, h0, _, z, g = _gated_inputs(rng, B, T, C, dtype)
    r_np, i_np = ri[..., :C], ri[..., C:]
    h_ref = O.gated_scan_fwd(xp, r_np, i_np, lam, h0)
    y_ref = O.silu(z) * h_ref
    dh = g * O.silu(z)
    dz_ref = g * h_ref * O.silu_grad(z)
    dxp, dr, di, dlam, dh0 = O.gated_scan_bwd(xp, r_np, i_np, lam, h0, dh)
    tol = TOL if dtype == torch.float32 else TOL_BF16
    txp, tri, tz = cuda(xp, dtype, True), cuda(ri, dtype, True), cuda(z, dtype, True)
    tlam, th0 = cuda(lam, torch.float32, True), cuda(h0, torch.float32, True)
    r_t, i_t = tri.chunk(2, dim=-1)
    y = _ops().gated_scan(txp, r_t, i_t, tlam, th0, z=tz)
    y.backward(cuda(g, dtype))
    assert rel_err(y, y_ref) <= tol
    # with bf16 storage dz/dx see the bf16-rounded h the kernel saved, hence the looser bound
    assert rel_err(tz.grad, dz_ref) <= tol
    assert rel_err(txp.grad, dxp) <= tol
    assert rel_err(tri.grad[..., :C], dr) <= tol
    assert rel_err(tri.grad[..., C:], di) <= tol
    assert rel_err(th0.grad, dh0) <= tol
    assert rel_err(tlam.grad, dlam) <= (5e-4 if dtype == torch.float32 else 5e-2)


@pytest.mark.parametrize("B,T,C", [(2, 9, 8), (2, 50, 128), (2, 300, 64)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_scan_channel_last(B, T, C, dtype):
    rng = np.random.default_rng(T)
    a = rng.uniform(0.6, 1.0, (B, T, C))
    b = rng.normal(size=(B, T, C))
    g = rng.normal(size=(B, T, C))
    h0 = rng.normal(size=C).astype(np.float32).astype(np.float64)
    if dtype == torch.bfloat16:
        rd = lambda v: torch.tensor(v, dtype=torch.float32).to(torch.bfloat16).double().numpy()
        a, b, g = rd(a), rd(b), rd(g)
    sw = lambda v: np.swapaxes(v, 1, 2)
    h_ref = O.scan_fwd(sw(a), sw(b), np.broadcast_to(h0, (B, C)).copy())
    da, db, dh0 = O.scan_bwd(sw(a), h_ref, sw(g), np.broadcast_to(h0, (B, C)))
    ta, tb, th0 = cuda(a, dtype, True), cuda(b, dtype, True), cuda(h0, torch.float32, True)
    h = _ops().scan_channel_last(ta, tb, th0)
    h.backward(cuda(g, dtype))
    tol = TOL if dtype == torch.float32 else TOL_BF16
    assert rel_err(h, sw(h_ref)) <= tol
    assert rel_err(ta.grad, sw(da)) <= tol
    assert rel_err(tb.grad, sw(db)) <= tol
    assert rel_err(th0.grad, dh0.sum(0)) <= tol


@pytest.mark.parametrize("B,T,C,W", [(2, 1, 4, 4), (2, 5, 8, 4), (3, 50, 128, 4), (2, 200, 128, 4), (2, 37, 64, 3),
                                     (2, 19, 12, 2), (64, 50, 128, 4), (2, 1000, 256, 4)])
@pytest.mark.parametrize("silu,bias", [(True, True), (False, True), (True, False)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_causal_conv1d(B, T, C, W, silu, bias, dtype):
    rng = np.random.default_rng(B + T + C + W)
    xz = rng.normal(size=(B, T, 2 * C))
    w = rng.normal(size=(C, W)) * 0.5
    bv = rng.normal(size=C) * 0.5 if bias else None
    g = rng.normal(size=(B, T, C))
    if dtype == torch.bfloat16:
        rd = lambda v: torch.tensor(v, dtype=torch.float32).to(torch.bfloat16).double().numpy()
        xz, g = rd(xz), rd(g)
    w = w.astype(np.float32).astype(np.float64)
    bv = bv.astype(np.float32).astype(np.float64) if bias else None
    x = xz[..., :C]
    y_ref = O.causal_conv1d_silu_fwd(x, w, bv, activation=silu)
    dx_ref, dw_ref, db_ref = O.causal_conv1d_silu_bwd(x, w, bv, g, activation=silu)
    txz = cuda(xz, dtype, True)
    tw = cuda(w, torch.float32, True)
    tb = cuda(bv, torch.float32, True) if bias else None
    tx = txz.chunk(2, dim=-1)[0]       # strided view, like xz.chunk(2, -1)[0] in RecBLR.py:174
    y = _ops().causal_conv1d_channel_last(tx, tw, tb, silu=silu)
    y.backward(cuda(g, dtype))
    tol = TOL if dtype == torch.float32 else TOL_BF16
    assert rel_err(y, y_ref) <= tol
    assert rel_err(txz.grad[..., :C], dx_ref) <= tol
    assert rel_err(tw.grad, dw_ref) <= (5e-4 if dtype == torch.float32 else 5e-2)
    if bias:
        assert rel_err(tb.grad, db_ref) <= (5e-4 if dtype == torch.float32 else 5e-2)


def test_causal_conv1d_fn_dropin_signature():
    """Called exactly like RecBLR.py:188-193: x = [B, C, T] view with channel-last strides."""
    from datamining_recblr_b200.causal_conv1d import causal_conv1d_fn
    rng = np.random.default_rng(0)
    x = rng.normal(size=(2, 20, 16))
    w = rng.normal(size=(16, 4))
    b = rng.normal(size=16)
    tx = cuda(x)
    out = causal_conv1d_fn(x=tx.mT, weight=cuda(w), bias=cuda(b), activation="silu").mT
    assert rel_err(out, O.causal_conv1d_silu_fwd(x, w.astype(np.float32).astype(np.float64),
                                                b.astype(np.float32).astype(np.float64))) <= TOL
    with pytest.raises(NotImplementedError):
        causal_conv1d_fn(x=tx.mT, weight=cuda(w), bias=cuda(b), activation="relu")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gated_scan_packed_equals_unpacked(dtype):
    """gated_scan_packed(x, ri) == gated_scan(x, *ri.chunk(2)) bit for bit, forward and every gradient."""
    rng = np.random.default_rng(3)
    B, T, C = 3, 77, 128
    xp, ri, lam, h0, _, z, g = _gated_inputs(rng, B, T, C, dtype)
    outs = []
    for packed in (False, True):
        txp, tri, tz = cuda(xp, dtype, True), cuda(ri, dtype, True), cuda(z, dtype, True)
        tlam, th0 = cuda(lam, torch.float32, True), cuda(h0, torch.float32, True)
        if packed:
            y = _ops().gated_scan_packed(txp, tri, tlam, th0, z=tz)
        else:
            r_t, i_t = tri.chunk(2, dim=-1)
            y = _ops().gated_scan(txp, r_t, i_t, tlam, th0, z=tz)
        y.backward(cuda(g, dtype))
        outs.append((y.detach(), txp.grad, tri.grad, tz.grad, tlam.grad, th0.grad))
    for a, b in zip(*outs):
        assert torch.equal(a, b)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gated_scan_saturated_gates_stay_finite(dtype):
    """Saturated recurrence gates (sigmoid(r) == 0 -> alpha == 1 exactly): sqrt(1 - alpha^2 + 1e-8) must stay 1e-4, never
    rsqrt(0) (regression: the bf16 one-FMA form once folded the 1e-8 into the FMA constant, where it rounds away)."""
    B, T, C = 3, 21, 64
    g = torch.Generator(device="cuda").manual_seed(0)
    xp = torch.randn(B, T, C, device="cuda", generator=g).to(dtype).requires_grad_(True)
    ri = torch.randn(B, T, 2 * C, device="cuda", generator=g)
    ri[..., :C] = -60.0                       # sigmoid(r) underflows to 0
    ri = ri.to(dtype).requires_grad_(True)
    lam = torch.linspace(-2.2, -6.9, C, device="cuda").requires_grad_(True)
    z = torch.randn(B, T, C, device="cuda", generator=g).to(dtype).requires_grad_(True)
    y = _ops().gated_scan_packed(xp, ri, lam, z=z)
    y.sum().backward()
    for t in (y, xp.grad, ri.grad, z.grad, lam.grad):
        assert torch.isfinite(t).all()
    # alpha == 1, beta' = 1e-4 * sigmoid(i) * x': h is a plain cumulative sum of tiny terms
    h_ref = (1e-4 * torch.sigmoid(ri[..., C:].double()) * xp.double()).cumsum(1) * torch.nn.functional.silu(z.double())
    assert float((y.double() - h_ref).abs().max()) <= (1e-6 if dtype == torch.float32 else 2e-2 * float(h_ref.abs().max()) + 1e-6)
