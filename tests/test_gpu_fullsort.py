"""GPU parity of the tcgen05 full-sort kernels (through the C ABI) against the numpy oracle fed the SAME bf16-rounded
operands: top-k ids identical under the lowest-index tie-break; CE statistics within 1e-5 relative."""
import numpy as np
import pytest
import torch

from oracle import bdlru_oracle as O

pytestmark = pytest.mark.gpu


def _bf(a):
    return torch.tensor(a, dtype=torch.float32).to(torch.bfloat16)


def _oracle_topk(qb, eb, k, mask_id):
    s = qb.double().numpy() @ eb.double().numpy().T
    if mask_id >= 0:
        s[:, mask_id] = -np.inf
    order = np.argsort(-s, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(s, order, axis=1), order, s


@pytest.mark.parametrize("B,N,D,k", [(128, 128, 64, 10), (256, 1000, 64, 10), (300, 3417, 64, 20), (4096, 12102, 64, 10),
                                      (130, 5000, 128, 10), (512, 20000, 128, 32), (64, 777, 256, 10), (200, 4097, 192, 5),
                                      (1, 50, 64, 10), (7, 9, 64, 10)])
def test_topk_random_matches_oracle(B, N, D, k):
    from datamining_recblr_b200 import ops
    rng = np.random.default_rng(B + N + D)
    qb, eb = _bf(rng.normal(size=(B, D))), _bf(rng.normal(size=(N, D)) * 0.02)
    vals, ids = ops.fullsort_topk(qb.cuda(), eb.cuda(), k, mask_id=0)
    v_ref, i_ref, s = _oracle_topk(qb, eb, k, 0)
    ids, vals = ids.cpu().numpy(), vals.cpu().numpy()
    kk = min(k, N - 1)
    # fp32 tensor-core accumulation vs float64: ids must agree except where two candidates are closer than fp32
    # accumulation noise (documented); with random data that never happens at these sizes
    bad = ids[:, :kk] != i_ref[:, :kk]
    if bad.any():
        rows = np.where(bad.any(axis=1))[0]
        for r in rows:
            a, b = s[r, ids[r, :kk]], s[r, i_ref[r, :kk]]
            assert np.abs(a - b).max() <= 1e-6 * np.abs(b).max(), (r, ids[r], i_ref[r])
    assert np.abs(vals[:, :kk] - v_ref[:, :kk]).max() <= 1e-5 * max(np.abs(v_ref[:, :kk]).max(), 1e-30)
    if kk < k:  # fewer than k candidates: (-inf, -1) padding
        assert (ids[:, kk:] == -1).all() and np.isneginf(vals[:, kk:]).all()


@pytest.mark.parametrize("D", [64, 128])
def test_topk_exact_ties_lowest_index_first(D):
    """Small-integer operands make every dot product exact in fp32, with massive exact ties: the kernel must order
    ties by lowest item id, exactly like a stable descending sort (north_star tie-break)."""
    from datamining_recblr_b200 import ops
    rng = np.random.default_rng(D)
    B, N, k = 260, 3000, 20
    q = rng.integers(-2, 3, size=(B, D)).astype(np.float64)
    e = rng.integers(-1, 2, size=(N, D)).astype(np.float64)
    e[1500:] = e[:1500]   # every item has an exact duplicate 1500 ids later
    qb, eb = _bf(q), _bf(e)
    vals, ids = ops.fullsort_topk(qb.cuda(), eb.cuda(), k, mask_id=0)
    v_ref, i_ref, _ = _oracle_topk(qb, eb, k, 0)
    assert (ids.cpu().numpy() == i_ref).all()
    assert (vals.cpu().numpy() == v_ref).all()


def test_topk_shard_offsets_and_merge_equal_single_table():
    """Row-sharded table: per-shard top-k with id_offset + bdlru_topk_merge == top-k over the whole table; the masked
    item (global id 0) lives in shard 0 only."""
    from datamining_recblr_b200 import ops
    rng = np.random.default_rng(5)
    B, N, D, k, G = 300, 9001, 64, 10, 4
    q = rng.integers(-2, 3, size=(B, D)).astype(np.float64)
    e = rng.integers(-1, 2, size=(N, D)).astype(np.float64)
    qb, eb = _bf(q).cuda(), _bf(e).cuda()
    v1, i1 = ops.fullsort_topk(qb, eb, k, mask_id=0)
    bounds = [N * g // G for g in range(G + 1)]
    cs, ci = [], []
    for g in range(G):
        v, i = ops.fullsort_topk(qb, eb[bounds[g]:bounds[g + 1]], k, mask_id=0, id_offset=bounds[g])
        cs.append(v)
        ci.append(i)
    vm, im = ops.topk_merge(torch.cat(cs, 1), torch.cat(ci, 1), k)
    assert torch.equal(im, i1) and torch.equal(vm, v1)
    _, i_ref, _ = _oracle_topk(qb.cpu(), eb.cpu(), k, 0)
    assert (i1.cpu().numpy() == i_ref).all()


def test_topk_no_mask_and_mask_other_id():
    from datamining_recblr_b200 import ops
    rng = np.random.default_rng(9)
    qb, eb = _bf(rng.normal(size=(140, 64))), _bf(rng.normal(size=(700, 64)))
    for mask in (-1, 333):
        _, ids = ops.fullsort_topk(qb.cuda(), eb.cuda(), 10, mask_id=mask)
        _, i_ref, _ = _oracle_topk(qb, eb, 10, mask)
        assert (ids.cpu().numpy() == i_ref).all()


@pytest.mark.parametrize("B,N,D", [(128, 128, 64), (300, 3417, 64), (2048, 12102, 64), (130, 5000, 128), (64, 777, 256),
                                    (5, 40, 64)])
def test_ce_stats_match_oracle(B, N, D):
    from datamining_recblr_b200 import ops
    rng = np.random.default_rng(B + N)
    qb, eb = _bf(rng.normal(size=(B, D)) * 2), _bf(rng.normal(size=(N, D)) * 0.3)
    pos = rng.integers(0, N, size=B)
    pos[0] = 0           # the pad row is a legal class
    m, s, pl = ops.fullsort_ce_stats(qb.cuda(), eb.cuda(), torch.tensor(pos).cuda())
    logits = qb.double().numpy() @ eb.double().numpy().T
    lse_ref = np.log(np.exp(logits - logits.max(1, keepdims=True)).sum(1)) + logits.max(1)
    lse = (m.double() + torch.log(s.double())).cpu().numpy()
    assert np.abs(lse - lse_ref).max() <= 1e-5 * np.abs(lse_ref).max()
    assert np.abs(pl.cpu().numpy() - logits[np.arange(B), pos]).max() <= 1e-5 * np.abs(logits).max()
    loss_ref, _, _, _ = O.ce_loss(qb.double().numpy(), eb.double().numpy(), pos)
    loss = float((m.double() + torch.log(s.double()) - pl.double()).mean())
    assert abs(loss - loss_ref) <= 1e-5 * abs(loss_ref)


@pytest.mark.parametrize("B,N,D", [(128, 128, 64), (300, 3417, 64), (2048, 12102, 64), (130, 5000, 128), (64, 777, 256),
                                    (200, 1000, 192), (5, 40, 64), (1000, 200, 128)])
def test_fused_ce_forward_backward_match_oracle(B, N, D):
    """Fused CE (tcgen05 forward statistics + two recompute-GEMM gradient passes) vs the float64 oracle on the same
    bf16-rounded operands.  Probabilities are rounded to bf16 before the gradient GEMMs (as in any flash-style
    backward), hence the 1e-2 bound on the gradients; the loss itself is at fp32 accuracy."""
    from datamining_recblr_b200 import ops
    rng = np.random.default_rng(B * 7 + N)
    qb, eb = _bf(rng.normal(size=(B, D)) * 2), _bf(rng.normal(size=(N, D)) * 0.3)
    pos = rng.integers(0, N, size=B)
    pos[0] = 0
    q = qb.float().cuda().requires_grad_(True)
    e = eb.float().cuda().requires_grad_(True)
    loss = ops.fullsort_cross_entropy(q, e, torch.tensor(pos).cuda())
    (loss * 3.0).backward()
    loss_ref, _, dQ, dE = O.ce_loss(qb.double().numpy(), eb.double().numpy(), pos)
    assert abs(float(loss) - loss_ref) <= 1e-5 * abs(loss_ref)
    gq, ge = q.grad.double().cpu().numpy() / 3.0, e.grad.double().cpu().numpy() / 3.0
    assert np.abs(gq - dQ).max() <= 1e-2 * np.abs(dQ).max()
    assert np.abs(ge - dE).max() <= 1e-2 * np.abs(dE).max()
    assert np.abs(ge[0]).max() > 0   # the pad row receives gradient as a never/rarely-positive class (quirk 2)


# ----------------------------------------------------------------------------- item bias (BERT4Rec scorer)
@pytest.mark.parametrize("B,N,D,k", [(256, 1000, 64, 10), (130, 5000, 128, 20), (64, 777, 192, 10)])
def test_topk_with_item_bias_matches_oracle(B, N, D, k):
    """scores = Q E^T + output_bias (bert4rec.py:230-242): the fp32 bias rides the GEMM as an exact 3-way bf16 split."""
    from datamining_recblr_b200 import ops
    rng = np.random.default_rng(B + N + D + 1)
    qb, eb = _bf(rng.normal(size=(B, D))), _bf(rng.normal(size=(N, D)) * 0.05)
    bias = torch.tensor(rng.normal(size=N) * 0.5, dtype=torch.float32)   # fp32, NOT bf16-representable
    vals, ids = ops.fullsort_topk(qb.cuda(), eb.cuda(), k, mask_id=0, item_bias=bias.cuda())
    s = O.full_sort_scores(qb.double().numpy(), eb.double().numpy(), bias.double().numpy())
    v_ref, i_ref = O.topk_lowest_index(s, k)
    ids, vals = ids.cpu().numpy(), vals.cpu().numpy()
    bad = ids != i_ref
    for r in np.where(bad.any(axis=1))[0]:   # only fp32-accumulation near-ties may swap
        assert np.abs(s[r, ids[r]] - s[r, i_ref[r]]).max() <= 1e-6 * np.abs(s[r, i_ref[r]]).max(), (r, ids[r], i_ref[r])
    assert np.abs(vals - v_ref).max() <= 1e-5 * np.abs(v_ref).max()
    # the bias must matter: without it the ranking differs
    _, ids0 = ops.fullsort_topk(qb.cuda(), eb.cuda(), k, mask_id=0)
    assert (ids0.cpu().numpy() != i_ref).any()


def test_topk_with_integer_bias_exact_ties():
    """Integer operands and integer biases: every score exact, massive ties, lowest id first."""
    from datamining_recblr_b200 import ops
    rng = np.random.default_rng(11)
    B, N, D, k = 200, 3000, 64, 20
    q = rng.integers(-2, 3, size=(B, D)).astype(np.float64)
    e = rng.integers(-1, 2, size=(N, D)).astype(np.float64)
    bias = rng.integers(-3, 4, size=N).astype(np.float64)
    e[1500:], bias[1500:] = e[:1500], bias[:1500]
    vals, ids = ops.fullsort_topk(_bf(q).cuda(), _bf(e).cuda(), k, mask_id=0, item_bias=torch.tensor(bias).float().cuda())
    v_ref, i_ref = O.topk_lowest_index(O.full_sort_scores(q, e, bias), k)
    assert (ids.cpu().numpy() == i_ref).all()
    assert (vals.cpu().numpy() == v_ref).all()


@pytest.mark.parametrize("B,N,D", [(300, 3417, 64), (130, 5000, 128), (64, 777, 192)])
def test_fused_ce_with_item_bias_forward_backward(B, N, D):
    """CE over logits Q E^T + bias (bert4rec.py:200-213): loss at fp32 accuracy, dQ / dE / d(bias) within the bf16-P bound."""
    from datamining_recblr_b200 import ops
    rng = np.random.default_rng(B * 3 + N)
    qb, eb = _bf(rng.normal(size=(B, D)) * 2), _bf(rng.normal(size=(N, D)) * 0.3)
    bias = rng.normal(size=N).astype(np.float32)
    pos = rng.integers(0, N, size=B)
    q = qb.float().cuda().requires_grad_(True)
    e = eb.float().cuda().requires_grad_(True)
    bt = torch.tensor(bias).cuda().requires_grad_(True)
    loss = ops.fullsort_cross_entropy(q, e, torch.tensor(pos).cuda(), item_bias=bt)
    (loss * 2.0).backward()
    b64 = bias.astype(np.float64)
    loss_ref, _, dQ, dE = O.ce_loss(qb.double().numpy(), eb.double().numpy(), pos, b64)
    db = O.ce_bias_grad(qb.double().numpy(), eb.double().numpy(), pos, b64)
    assert abs(float(loss) - loss_ref) <= 1e-5 * abs(loss_ref)
    assert q.grad.shape == (B, D) and e.grad.shape == (N, D) and bt.grad.shape == (N,)
    assert np.abs(q.grad.double().cpu().numpy() / 2 - dQ).max() <= 1e-2 * np.abs(dQ).max()
    assert np.abs(e.grad.double().cpu().numpy() / 2 - dE).max() <= 1e-2 * np.abs(dE).max()
    assert np.abs(bt.grad.double().cpu().numpy() / 2 - db).max() <= 1e-2 * np.abs(db).max()


def test_baseline_model_heads_match_reference_expressions():
    """scoring.py against the literal torch expressions of bert4rec.py:200-213,230-242 (bias, [MASK] row dropped,
    target-weighted CE) and sasrec.py:129-133 on the same bf16-rounded operands."""
    from datamining_recblr_b200 import scoring
    g = torch.Generator(device="cuda").manual_seed(4)
    B, ML, H, n_items = 48, 5, 64, 900
    w = (torch.randn(n_items + 1, H, device="cuda", generator=g) * 0.3).bfloat16().float().requires_grad_(True)
    bias = torch.randn(n_items, device="cuda", generator=g).requires_grad_(True)
    seq = torch.randn(B, ML, H, device="cuda", generator=g).bfloat16().float().requires_grad_(True)
    pos = torch.randint(0, n_items, (B, ML), device="cuda", generator=g)
    masked_index = torch.randint(0, 3, (B, ML), device="cuda", generator=g)
    # reference, in float64
    w64, b64, s64 = (t.detach().double().requires_grad_(True) for t in (w, bias, seq))
    logits = torch.matmul(s64, w64[:n_items].transpose(0, 1)) + b64
    targets = (masked_index > 0).double().view(-1)
    ref = torch.sum(torch.nn.CrossEntropyLoss(reduction="none")(logits.view(-1, n_items), pos.view(-1)) * targets) / targets.sum()
    ref.backward()
    loss = scoring.cross_entropy(seq, w, pos, n_items=n_items, output_bias=bias, targets=masked_index > 0)
    loss.backward()
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    for got, want in ((seq.grad, s64.grad), (w.grad, w64.grad), (bias.grad, b64.grad)):
        assert got.shape == want.shape
        assert (got.double() - want).abs().max() <= 1e-2 * want.abs().max()
    assert w.grad[n_items].abs().max() == 0       # the [MASK] row takes no part in the softmax
    # BERT4Rec full sort (last position) and SASRec (no bias, whole table)
    q = seq[:, -1].detach()
    sc, ids = scoring.full_sort_topk(q, w.detach(), 10, n_items=n_items, output_bias=bias.detach())
    dense = (q.double() @ w64[:n_items].detach().T + b64.detach()).cpu().numpy()
    v_ref, i_ref = O.topk_lowest_index(dense, 10)
    assert (ids.cpu().numpy() == i_ref).all() and ids.dtype == torch.int64
    sc2, ids2 = scoring.full_sort_topk(q, w.detach(), 10)
    v2, i2 = O.topk_lowest_index((q.double() @ w64.detach().T).cpu().numpy(), 10)
    assert (ids2.cpu().numpy() == i2).all()
    l2 = scoring.cross_entropy(q, w.detach(), pos[:, 0])
    ref2 = torch.nn.functional.cross_entropy(q.double() @ w64.detach().T, pos[:, 0])
    assert abs(float(l2) - float(ref2)) <= 1e-5 * abs(float(ref2))


@pytest.mark.parametrize("with_bias", [False, True])
def test_sharded_ce_wrapper_single_rank_matches_fused_ce(with_bias):
    """sharded.sharded_cross_entropy without a process group (world 1) = ops.fullsort_cross_entropy: same loss and
    gradients (the multi-rank exchange itself is covered by the gloo tests and tests/dist_check.py)."""
    from datamining_recblr_b200 import ops, sharded
    g = torch.Generator(device="cuda").manual_seed(9)
    B, N, D = 200, 2500, 64
    q0 = torch.randn(B, D, device="cuda", generator=g)
    e0 = torch.randn(N, D, device="cuda", generator=g) * 0.3
    b0 = torch.randn(N, device="cuda", generator=g) if with_bias else None
    pos = torch.randint(0, N, (B,), device="cuda", generator=g)
    outs = []
    for fn in (lambda q, e, b: ops.fullsort_cross_entropy(q, e, pos, item_bias=b),
               lambda q, e, b: sharded.sharded_cross_entropy(q, e, pos, id_offset=0, item_bias=b)):
        q, e = q0.clone().requires_grad_(True), e0.clone().requires_grad_(True)
        b = b0.clone().requires_grad_(True) if with_bias else None
        loss = fn(q, e, b)
        loss.backward()
        outs.append((loss.detach(), q.grad, e.grad, None if b is None else b.grad))
    # fullsort_cross_entropy takes the fused forward (dQ out of the forward pass), the sharded wrapper the statistics
    # kernel + both gradient passes: same mathematics, different bf16 rounding points of P
    assert abs(float(outs[0][0]) - float(outs[1][0])) <= 1e-6 * abs(float(outs[0][0]))
    for a, c in zip(outs[0][1:], outs[1][1:]):
        assert (a is None and c is None) or float((a - c).abs().max()) <= 1e-2 * float(a.abs().max())


def test_item_bias_unsupported_width_raises():
    from datamining_recblr_b200 import ops
    assert ops.fullsort_bias_supported(64) and ops.fullsort_bias_supported(192) and not ops.fullsort_bias_supported(256)


def test_sharded_paths_on_two_gpus():
    """Launches tests/dist_check.py on 2 GPUs of this node (NCCL); skipped on a 1-GPU box."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631", os.path.join(root, "tests", "dist_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "dist_check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_ce_stats_shard_with_positive_outside_tail_tile():
    """Regression: a user whose positive item lives in ANOTHER shard, at a global id that falls inside this shard's last
    (partial, masked) tile, must keep pos_logit = 0 instead of picking up the -inf of a masked tail row."""
    from datamining_recblr_b200 import ops
    rng = np.random.default_rng(1)
    B, N, D = 64, 1500, 64          # 1500 = 15 * 96 + 60: the last tile covers ids 1440..1535
    qb, eb = _bf(rng.normal(size=(B, D))), _bf(rng.normal(size=(N, D)) * 0.1)
    pos = np.full(B, 1510)          # outside [0, 1500) but inside the tail tile's id range
    pos[:5] = [0, 7, 1499, 1441, 1500]
    m, s, pl = ops.fullsort_ce_stats(qb.cuda(), eb.cuda(), torch.tensor(pos).cuda())
    logits = qb.double().numpy() @ eb.double().numpy().T
    assert torch.isfinite(pl).all()
    assert np.abs(pl[:4].cpu().numpy() - logits[np.arange(4), pos[:4]]).max() <= 1e-5 * np.abs(logits).max()
    assert float(pl[4:].abs().max()) == 0.0


@pytest.mark.parametrize("B,N,D,k", [(1, 5, 64, 32), (3, 40, 128, 32), (129, 97, 64, 1), (257, 193, 256, 20)])
def test_topk_edge_sizes(B, N, D, k):
    """Single user, k larger than the catalog (padding with (-inf, -1)), k = 1, sizes one past a tile / block boundary."""
    from datamining_recblr_b200 import ops
    rng = np.random.default_rng(B * N)
    qb, eb = _bf(rng.integers(-2, 3, size=(B, D))), _bf(rng.integers(-1, 2, size=(N, D)))
    vals, ids = ops.fullsort_topk(qb.cuda(), eb.cuda(), k, mask_id=0)
    v_ref, i_ref, _ = _oracle_topk(qb, eb, k, 0)
    kk = min(k, N - 1)
    assert (ids.cpu().numpy()[:, :kk] == i_ref[:, :kk]).all()
    assert (vals.cpu().numpy()[:, :kk] == v_ref[:, :kk]).all()
    if kk < k:
        assert (ids.cpu().numpy()[:, kk:] == -1).all()


def test_fullsort_rejects_bad_arguments():
    from datamining_recblr_b200 import ops
    from datamining_recblr_b200._lib import BdlruError
    q = torch.randn(4, 48, device="cuda")
    e = torch.randn(10, 48, device="cuda")
    with pytest.raises(BdlruError):          # D must be a multiple of 64
        ops.fullsort_topk(q, e, 5)
    q = torch.randn(4, 64, device="cuda")
    e = torch.randn(10, 64, device="cuda")
    with pytest.raises(BdlruError):          # k <= 32
        ops.fullsort_topk(q, e, 33)
    with pytest.raises(Exception):           # no CPU fallback
        ops.fullsort_topk(q.cpu(), e.cpu(), 5)


@pytest.mark.parametrize("B,N,D", [(200, 40_000, 128), (130, 19_001, 64), (96, 37_999, 128), (300, 19_073, 192)])
def test_ce_de_pass_row_block_handover(B, N, D):
    """dE pass where every CTA hands over between several 128-row blocks (N > 148 * 128 item rows) and the last block is
    ragged: the next block's X rows arrive by TMA into the stage the drained dX leaves through (bulk tensor stores clipped
    at N), D = 64 shares one X slab between the two softmax groups, D = 192 takes the path without the stage.  Reference:
    float64 softmax gradients of the same bf16 operands (RecBLR.py:99-103 through autograd)."""
    from datamining_recblr_b200 import ops
    torch.manual_seed(B + N)
    q = torch.randn(B, D, device="cuda").to(torch.bfloat16)
    e = (torch.randn(N, D, device="cuda") * 0.2).to(torch.bfloat16)
    pos = torch.randint(0, N, (B,), device="cuda")
    pos[0], pos[1] = N - 1, 0                      # positives in the ragged last block and in the first row
    m, s, _ = ops.fullsort_ce_stats(q, e, pos)
    lse = m + torch.log(s)
    de_guard = torch.full((N + 128, D), 7.0, device="cuda")     # rows past N must stay untouched by the bulk stores
    dq, de = ops.fullsort_ce_grads(q, e, pos, lse, 1.0 / B, out_de=de_guard[:N])
    assert torch.all(de_guard[N:] == 7.0)
    p64 = torch.softmax(q.double() @ e.double().T, -1)
    p64[torch.arange(B, device="cuda"), pos] -= 1.0
    de64, dq64 = p64.T @ q.double() / B, p64 @ e.double() / B
    assert ((de.double() - de64).abs().max() / de64.abs().max()).item() <= 1e-2
    assert ((dq.double() - dq64).abs().max() / dq64.abs().max()).item() <= 1e-2


@pytest.mark.parametrize("B,N,D", [(200, 20000, 64), (20000, 300, 64), (19000, 19500, 128)])
def test_fused_ce_many_row_blocks(B, N, D):
    """More row blocks than SMs in the dE pass (N > 148 * 128) and in the dQ pass (B > 148 * 128): every CTA of the
    persistent backward kernel walks several work items (X reloaded into TMEM, accumulator reused), and the forward /
    top-k kernels run several waves of user groups."""
    from datamining_recblr_b200 import ops
    rng = np.random.default_rng(B + N)
    qb, eb = _bf(rng.normal(size=(B, D)) * 1.5), _bf(rng.normal(size=(N, D)) * 0.3)
    pos = rng.integers(0, N, size=B)
    q = qb.float().cuda().requires_grad_(True)
    e = eb.float().cuda().requires_grad_(True)
    loss = ops.fullsort_cross_entropy(q, e, torch.tensor(pos).cuda())
    loss.backward()
    qd, ed = qb.double().cuda(), eb.double().cuda()     # float64 reference on the GPU (the matrices are large)
    logits = qd @ ed.T
    lse = torch.logsumexp(logits, dim=1)
    tp = torch.tensor(pos).cuda()
    loss_ref = (lse - logits[torch.arange(B), tp]).mean()
    p = torch.exp(logits - lse[:, None])
    p[torch.arange(B), tp] -= 1.0
    p /= B
    dQ, dE = p @ ed, p.T @ qd
    assert abs(float(loss) - float(loss_ref)) <= 1e-5 * abs(float(loss_ref))
    assert float((q.grad.double() - dQ).abs().max()) <= 1e-2 * float(dQ.abs().max())
    assert float((e.grad.double() - dE).abs().max()) <= 1e-2 * float(dE.abs().max())
    vals, ids = ops.fullsort_topk(qb.cuda(), eb.cuda(), 10, mask_id=0)
    logits[:, 0] = float("-inf")
    ref_ids = torch.sort(logits, dim=1, descending=True, stable=True).indices[:, :10]
    agree = (ids.long() == ref_ids).all(dim=1).double().mean()
    assert float(agree) >= 0.999     # fp32 vs float64 near-ties only


@pytest.mark.parametrize("B,N,D", [(130, 5000, 128), (300, 3417, 64), (64, 777, 256), (2048, 40000, 128)])
def test_rowmax_exact_and_sampled(B, N, D):
    """bdlru_fullsort_rowmax: exact row maxima (tile_stride 1) == max of the float64 logits of the same bf16 operands;
    with tile_stride s it is exactly the maximum over the sampled 96-item tiles (never above the true maximum)."""
    from datamining_recblr_b200 import ops
    rng = np.random.default_rng(N)
    qb, eb = _bf(rng.normal(size=(B, D))).cuda(), _bf(rng.normal(size=(N, D)) * 0.3).cuda()
    logits = qb.double() @ eb.double().T
    m = ops.fullsort_rowmax(qb, eb, 1)
    ref = logits.max(1).values
    assert float((m.double() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
    NT = 96 if D <= 128 else 64
    for s in (2, 7):
        ms = ops.fullsort_rowmax(qb, eb, s)
        cols = torch.cat([torch.arange(t * NT, min((t + 1) * NT, N)) for t in range(0, -(-N // NT), s)]).cuda()
        want = logits[:, cols].max(1).values
        assert float((ms.double() - want).abs().max()) <= 1e-5 * float(ref.abs().max())
        assert bool((ms.double() <= ref + 1e-5 * ref.abs().max()).all())


@pytest.mark.parametrize("B,N,D", [(130, 5000, 128), (300, 3417, 64), (64, 777, 256), (1000, 30000, 128)])
@pytest.mark.parametrize("shift", [0.0, -25.0, 10.0])
def test_ce_fwd_dq_is_shift_invariant_and_matches_float64(B, N, D, shift):
    """bdlru_fullsort_ce_fwd_dq with reference = row maximum + shift: lse = ref + log s and acc / s = softmax(l) E do not
    depend on the reference (the property the sampled maximum relies on), and match float64 on the same bf16 operands."""
    from datamining_recblr_b200 import ops
    rng = np.random.default_rng(B + N)
    qb, eb = _bf(rng.normal(size=(B, D)) * 1.5).cuda(), _bf(rng.normal(size=(N, D)) * 0.3).cuda()
    logits = qb.double() @ eb.double().T
    ref = (logits.max(1).values + shift).float()
    acc, s = ops.fullsort_ce_fwd_dq(qb, eb, ref)
    lse = ref.double() + torch.log(s.double())
    lse_ref = torch.logsumexp(logits, dim=1)
    assert float((lse - lse_ref).abs().max()) <= 1e-5 * float(lse_ref.abs().max())
    pe = torch.softmax(logits, dim=1) @ eb.double()
    got = acc.double() / s.double()[:, None]
    assert float((got - pe).abs().max()) <= 1e-2 * float(pe.abs().max())     # P is rounded to bf16 for the second GEMM


def test_fused_ce_paths_agree():
    """q needs a gradient -> fused forward (+ dE-only backward); q frozen -> statistics kernel + full backward: same loss,
    same table gradient; and the exact (stride 1) and sampled (stride 16) references give the same result."""
    from datamining_recblr_b200 import ops
    rng = np.random.default_rng(3)
    B, N, D = 500, 20000, 128
    qb, eb = _bf(rng.normal(size=(B, D)) * 1.5), _bf(rng.normal(size=(N, D)) * 0.3)
    pos = torch.tensor(rng.integers(0, N, size=B)).cuda()
    res = []
    for q_grad, stride in ((True, 16), (True, 1), (False, 16)):
        ops.CE_REFERENCE_STRIDE = stride
        q = qb.float().cuda().requires_grad_(q_grad)
        e = eb.float().cuda().requires_grad_(True)
        loss = ops.fullsort_cross_entropy(q, e, pos)
        (loss * 1.5).backward()
        res.append((float(loss), e.grad.clone(), q.grad.clone() if q_grad else None))
    ops.CE_REFERENCE_STRIDE = 16
    for r in res[1:]:
        assert abs(r[0] - res[0][0]) <= 1e-6 * abs(res[0][0])
        assert float((r[1] - res[0][1]).abs().max()) <= 1e-3 * float(res[0][1].abs().max())   # bf16 rounding points of P
    # dQ: P = exp(l - ref) is rounded to bf16 before the second GEMM, so a different reference moves the rounding points
    assert float((res[1][2] - res[0][2]).abs().max()) <= 1e-2 * float(res[0][2].abs().max())
