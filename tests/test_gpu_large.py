"""GPU parity AT THE SHAPES THE PERFORMANCE CLAIMS ARE MADE ON (What's weak #2): BASELINE.json configs[2]
(gate+scan sweep corners), configs[3] (4 096 users x 1 M / 10 M items) and configs[4] (CE at 8 192 x 1 M x 128).

The numpy oracle cannot hold these sizes, so the reference is the same float64 restatement evaluated ON THE GPU in
chunks (torch.float64, sequential scan loop, dense logits per item chunk) from the same (bf16-rounded) operands —
`oracle/bdlru_oracle.py` pins that restatement at small sizes (tests/test_gpu_scan.py compares the two on one case).
Bars: fp32 I/O max-norm relative error <= 1e-4 AND element-wise relative error (floor 1 % of the largest magnitude)
<= 1e-3; top-k ids identical, EXACTLY for integer-valued operands (every dot product exact in fp32, massive ties) and up
to fp32-vs-float64 near-ties (< 1e-6 relative score gap, counted and bounded) for random operands."""
import numpy as np
import pytest
import torch

from tests.util import elem_rel_err

pytestmark = pytest.mark.gpu


# ----------------------------------------------------------------------------- float64 references on the GPU
def gated_scan_ref64(xp, r, i, lam, h0=None, z=None):
    """RecBLR.py:197-200 (+206 with z) in float64 with a sequential scan, autograd-capable."""
    c = torch.nn.functional.softplus(lam)
    a = torch.exp(-c * torch.sigmoid(r))
    b = torch.sqrt(1 - a * a + 1e-8) * torch.sigmoid(i) * xp
    B, T, C = xp.shape
    s = torch.zeros(B, C, dtype=xp.dtype, device=xp.device) if h0 is None else h0.expand(B, C)
    hs = []
    for t in range(T):
        s = a[:, t] * s + b[:, t]
        hs.append(s)
    h = torch.stack(hs, 1)
    return h if z is None else torch.nn.functional.silu(z) * h


def _maxnorm(x, ref):
    return float((x.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def _elem(x, ref, floor=1e-2):
    den = torch.maximum(ref.abs(), floor * ref.abs().max())
    return float(((x.double() - ref).abs() / den).max())


@pytest.mark.parametrize("B,T,C", [(256, 4096, 256), (2048, 200, 128)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gated_scan_at_sweep_shapes(B, T, C, dtype):
    """configs[2] corner 256 x 4096 x 256 and the model shape 2 048 x 200 x 128, z-gated, broadcast h0, fwd + bwd."""
    from datamining_recblr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(B + T + C)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
    xp, ri, z, gy = rn(B, T, C), rn(B, T, 2 * C) * 1.5, rn(B, T, C), rn(B, T, C)
    lam = torch.linspace(-2.2, -6.9, C, device="cuda") + 0.3 * rn(C)      # RecBLR.py:153-158 range, perturbed
    h0 = rn(C)
    xp, ri, z, gy = (t.to(dtype) for t in (xp, ri, z, gy))                # the kernel and the reference see the same values
    txp, tri, tz = (t.clone().requires_grad_(True) for t in (xp, ri, z))
    tlam, th0 = lam.clone().requires_grad_(True), h0.clone().requires_grad_(True)
    y = ops.gated_scan_packed(txp, tri, tlam, h0=th0, z=tz)
    y.backward(gy)
    # float64 reference, batch-chunked to bound memory (autograd through the T-step loop)
    xd, rid, zd, gd = (t.double() for t in (xp, ri, z, gy))
    y_ref = torch.empty(B, T, C, dtype=torch.float64, device="cuda")
    dx_ref, dri_ref, dz_ref = torch.empty_like(y_ref), torch.empty(B, T, 2 * C, dtype=torch.float64, device="cuda"), \
        torch.empty_like(y_ref)
    dlam_ref, dh0_ref = torch.zeros(C, dtype=torch.float64, device="cuda"), torch.zeros(C, dtype=torch.float64, device="cuda")
    step = 32 if T > 1000 else 512
    for b0 in range(0, B, step):
        sl = slice(b0, b0 + step)
        a_, b_, c_ = (t[sl].clone().requires_grad_(True) for t in (xd, rid, zd))
        l_, h_ = lam.double().requires_grad_(True), h0.double().requires_grad_(True)
        yr = gated_scan_ref64(a_, b_[..., :C], b_[..., C:], l_, h_, c_)
        yr.backward(gd[sl])
        y_ref[sl], dx_ref[sl], dri_ref[sl], dz_ref[sl] = yr.detach(), a_.grad, b_.grad, c_.grad
        dlam_ref += l_.grad
        dh0_ref += h_.grad
    # max-norm: the north star's 1e-4 (measured 8e-6 at T = 4096, 1.5e-6 at T = 200).  Element-wise with a floor of 1 % of
    # the largest magnitude: 1e-3 — entries that small are sums that cancelled, so their error is set by the size of the
    # terms (measured 3.4e-4 at T = 4096, 7e-5 at T = 200; any fp32 scan, the reference's Triton kernel included, has this).
    if dtype == torch.float32:
        tol, etol, ptol = 1e-4, 1e-3, 1e-3
    else:   # bf16 I/O: one bf16 rounding per output (2^-9 relative) on top of the saved bf16 h the backward re-reads
        tol, etol, ptol = 2e-2, 5e-2, 5e-2
    errs = {"y": (y, y_ref), "dx'": (txp.grad, dx_ref), "dri": (tri.grad, dri_ref), "dz": (tz.grad, dz_ref)}
    rep = {name: (_maxnorm(got, want), _elem(got, want)) for name, (got, want) in errs.items()}
    rep["dLambda"] = (_maxnorm(tlam.grad, dlam_ref), None)
    rep["dh0"] = (_maxnorm(th0.grad, dh0_ref), None)
    print("gated_scan parity", (B, T, C), dtype, {k: tuple(None if x is None else float(f"{x:.3g}") for x in v) for k, v in rep.items()})
    for name in errs:
        assert rep[name][0] <= tol, (name, rep)
        assert rep[name][1] <= etol, (name, "element-wise", rep)
    assert rep["dLambda"][0] <= ptol and rep["dh0"][0] <= ptol, rep


def test_gated_scan_gpu_reference_equals_numpy_oracle():
    """The float64 GPU restatement used above == oracle/bdlru_oracle.py (the pinned oracle) on a small case."""
    from oracle import bdlru_oracle as O
    rng = np.random.default_rng(0)
    B, T, C = 3, 37, 8
    xp, r, i = rng.normal(size=(B, T, C)), rng.normal(size=(B, T, C)), rng.normal(size=(B, T, C))
    lam, h0 = O.lambda_init(C), rng.normal(size=C)
    t = lambda a: torch.tensor(a, device="cuda")
    h = gated_scan_ref64(t(xp), t(r), t(i), t(lam), t(h0))
    assert np.abs(h.cpu().numpy() - O.gated_scan_fwd(xp, r, i, lam, h0)).max() <= 1e-12


# ----------------------------------------------------------------------------- full-sort top-k at 1 M / 10 M items
def _topk_ref64(qd, eb, k, mask_id, chunk=1 << 16, integer=False):
    """Exact float64 top-k by (score desc, id asc) over item chunks: per-chunk torch.topk, then a stable lexicographic
    merge of the candidates.  qd float64 [U, D]; eb bf16 [N, D] (chunks are widened on the fly).  integer=True: the scores
    are small integers with huge tie groups, so the per-chunk selection runs on the unique key score * 2^25 - id (exact in
    float64) — torch.topk alone returns an arbitrary subset of a tie group."""
    U, N = qd.shape[0], eb.shape[0]
    best_s = torch.full((U, k), float("-inf"), dtype=torch.float64, device=qd.device)
    best_i = torch.full((U, k), -1, dtype=torch.int64, device=qd.device)
    for c0 in range(0, N, chunk):
        s = qd @ eb[c0:c0 + chunk].double().T
        if c0 <= mask_id < c0 + s.shape[1]:
            s[:, mask_id - c0] = float("-inf")
        kk = min(k, s.shape[1])
        # the k largest of the chunk INCLUDING all ties at the k-th value that matter: take 2k and let the merge decide
        if integer:
            key = s * float(1 << 25) - torch.arange(c0, c0 + s.shape[1], device=s.device, dtype=torch.float64)
            ci = torch.topk(key, kk, dim=1).indices
            cs = s.gather(1, ci)
        else:
            cs, ci = torch.topk(s, min(2 * kk, s.shape[1]), dim=1)
        cand_s = torch.cat([best_s, cs], 1)
        cand_i = torch.cat([best_i, ci + c0], 1)
        cand_i = torch.where(torch.isneginf(cand_s), torch.full_like(cand_i, 1 << 40), cand_i)
        o1 = torch.argsort(cand_i, dim=1, stable=True)                       # id ascending ...
        cand_s, cand_i = cand_s.gather(1, o1), cand_i.gather(1, o1)
        o2 = torch.argsort(cand_s, dim=1, descending=True, stable=True)      # ... then score descending, stable
        best_s, best_i = cand_s.gather(1, o2)[:, :k], cand_i.gather(1, o2)[:, :k]
    return best_s, best_i


def _check_topk(vals, ids, ref_s, ref_i, qd, eb, exact):
    ids = ids.long()
    same = ids == ref_i
    if exact:
        assert bool(same.all()), f"{int((~same).any(1).sum())} users differ"
        assert bool((vals.double() == ref_s).all())
        return 0
    rows = (~same).any(1).nonzero().flatten()
    for r in rows.tolist():   # only fp32-accumulation near-ties may differ: same SET up to swaps of near-equal scores
        got = (qd[r] @ eb[ids[r]].double().T)
        assert float((got - ref_s[r]).abs().max()) <= 1e-6 * float(ref_s[r].abs().max()), (r, ids[r], ref_i[r])
    assert float((vals.double() - ref_s).abs().max()) <= 1e-5 * float(ref_s.abs().max())
    return len(rows)


@pytest.mark.parametrize("k", [10, 20])
def test_topk_1m_items_4096_users(k):
    """configs[3] at 1 M items, all 4 096 users checked.  Random operands: ids identical up to near-ties (bounded);
    integer operands: every product exact in fp32 and ~10^5 exact ties per user -> ids must be EXACTLY the stable sort's,
    which also pins the `>=` threshold sharing between the 9 item splits."""
    from datamining_recblr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(2020)
    B, N, D = 4096, 1_000_000, 128
    eb = (torch.randn(N, D, device="cuda", generator=g) * 0.02).bfloat16()
    qb = torch.randn(B, D, device="cuda", generator=g).bfloat16()
    vals, ids = ops.fullsort_topk(qb, eb, k, mask_id=0)
    rs, ri = _topk_ref64(qb.double(), eb, k, 0)
    n_near = _check_topk(vals, ids, rs, ri, qb.double(), eb, exact=False)
    assert n_near <= B // 100, n_near
    ei = torch.randint(-1, 2, (N, D), device="cuda", generator=g).bfloat16()
    qi = torch.randint(-2, 3, (B, D), device="cuda", generator=g).bfloat16()
    vals, ids = ops.fullsort_topk(qi, ei, k, mask_id=0)
    rs, ri = _topk_ref64(qi.double(), ei, k, 0, integer=True)
    _check_topk(vals, ids, rs, ri, qi.double(), ei, exact=True)


def test_topk_10m_items_user_subsample():
    """configs[3] at 10 M items: the kernel scores all 4 096 users (the benchmarked launch); every 8th user is checked
    against the float64 reference."""
    from datamining_recblr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(7)
    B, N, D, k = 4096, 10_000_000, 128, 10
    eb = (torch.randn(N, D, device="cuda", generator=g) * 0.02).bfloat16()
    qb = torch.randn(B, D, device="cuda", generator=g).bfloat16()
    vals, ids = ops.fullsort_topk(qb, eb, k, mask_id=0)
    sub = torch.arange(0, B, 8, device="cuda")
    rs, ri = _topk_ref64(qb[sub].double(), eb, k, 0, chunk=1 << 18)
    n_near = _check_topk(vals[sub], ids[sub], rs, ri, qb[sub].double(), eb, exact=False)
    assert n_near <= 8, n_near
    # sharded form of the same launch (8 row shards, as 8 ranks would hold them) merged == single table
    bounds = [N * s // 8 for s in range(9)]
    parts = [ops.fullsort_topk(qb, eb[bounds[s]:bounds[s + 1]], k, mask_id=0, id_offset=bounds[s]) for s in range(8)]
    vm, im = ops.topk_merge(torch.cat([p[0] for p in parts], 1), torch.cat([p[1] for p in parts], 1), k)
    assert torch.equal(im, ids) and torch.equal(vm, vals)


# ----------------------------------------------------------------------------- CE at configs[4]'s per-GPU shape
def test_fused_ce_8192_users_1m_items():
    """CE fwd + bwd at 8 192 x 1 M x 128 (configs[4] per-GPU batch against a 1 M-row table / an 8-GPU shard of 8 M):
    loss <= 1e-5, dQ and dE <= 1e-2 max-norm (P is rounded to bf16 for the second GEMM) against float64."""
    from datamining_recblr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(11)
    B, N, D = 8192, 1_000_000, 128
    eb = (torch.randn(N, D, device="cuda", generator=g) * 0.05).bfloat16()
    qb = (torch.randn(B, D, device="cuda", generator=g) * 1.5).bfloat16()
    pos = torch.randint(0, N, (B,), device="cuda", generator=g)
    q = qb.float().requires_grad_(True)
    e = eb.float().requires_grad_(True)
    loss = ops.fullsort_cross_entropy(q, e, pos)
    loss.backward()
    qd = qb.double()
    chunk = 1 << 16
    m = torch.full((B,), float("-inf"), dtype=torch.float64, device="cuda")
    s = torch.zeros(B, dtype=torch.float64, device="cuda")
    for c0 in range(0, N, chunk):
        lg = qd @ eb[c0:c0 + chunk].double().T
        m2 = torch.maximum(m, lg.max(1).values)
        s = s * torch.exp(m - m2) + torch.exp(lg - m2[:, None]).sum(1)
        m = m2
    lse = m + torch.log(s)
    pl = (qd * eb[pos].double()).sum(1)
    loss_ref = (lse - pl).mean()
    assert abs(float(loss) - float(loss_ref)) <= 1e-5 * abs(float(loss_ref))
    dQ = torch.zeros(B, D, dtype=torch.float64, device="cuda")
    worst_e, scale_e = 0.0, 0.0
    ar = torch.arange(B, device="cuda")
    for c0 in range(0, N, chunk):
        ed = eb[c0:c0 + chunk].double()
        p = torch.exp(qd @ ed.T - lse[:, None])
        inside = (pos >= c0) & (pos < c0 + ed.shape[0])
        p[ar[inside], pos[inside] - c0] -= 1.0
        p /= B
        dQ += p @ ed
        dE = p.T @ qd
        worst_e = max(worst_e, float((e.grad[c0:c0 + chunk].double() - dE).abs().max()))
        scale_e = max(scale_e, float(dE.abs().max()))
    assert worst_e <= 1e-2 * scale_e, (worst_e, scale_e)
    assert float((q.grad.double() - dQ).abs().max()) <= 1e-2 * float(dQ.abs().max())


def test_elementwise_error_helper_is_stricter_than_maxnorm():
    ref = np.array([1.0, 1e-2, 1e-5])
    x = ref + np.array([0.0, 1e-5, 0.0])
    from tests.util import rel_err
    assert rel_err(x, ref) <= 1e-4 < elem_rel_err(x, ref)
