"""GPU parity against the REFERENCE'S OWN implementation launched on the B200 (What's missing #3): the
unmodified `parallel_scan.py` Triton kernels (parallel_scan.py:44-118, JIT-compiled for sm_100a by the image's triton)
and the unmodified `RecBLR.py` modules on top of them — literal left pad to a power of two, F.conv1d fallback (the
causal-conv1d wheel is absent), separate gate ops, two transposes, Triton scan.

The sources are not in the repository: `oracle/build_ref.py` packs them from /root/reference into the git-ignored
`oracle/_ref/reference_py.tar.gz`, which travels to the GPU box; without the blob these tests skip (and say so).
Tolerances: both sides compute in fp32, so each is also held to the float64 oracle; ours-vs-reference <= 1e-4 max-norm
(north_star) on outputs and input gradients."""
import numpy as np
import pytest
import torch

from oracle import build_ref
from tests.util import rel_err

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not build_ref.available(), reason="oracle/_ref/reference_py.tar.gz not built")]
TOL = 1e-4


@pytest.mark.parametrize("T", [64, 256, 1024, 4096])
@pytest.mark.parametrize("regime", ["model", "near_one"])
def test_parallel_scan_matches_reference_triton_kernel(T, regime):
    """ops.parallel_scan (csrc/scan_bct.cu) vs the reference's Triton `parallel_scan` on the same fp32 [B, C, T] inputs,
    forward and both gradients; and both against the float64 sequential scan."""
    from datamining_recblr_b200 import ops
    from oracle import bdlru_oracle as O
    _, ps = build_ref.load_gpu_reference()
    B, C = 6, 32
    g = torch.Generator(device="cuda").manual_seed(T)
    lo, hi = {"model": (0.3, 0.999), "near_one": (0.999, 1.0)}[regime]
    gates = torch.rand(B, C, T, device="cuda", generator=g) * (hi - lo) + lo
    tokens = torch.randn(B, C, T, device="cuda", generator=g)
    gy = torch.randn(B, C, T, device="cuda", generator=g)
    outs = []
    for fn in (ops.parallel_scan, ps.parallel_scan):
        a, b = gates.clone().requires_grad_(True), tokens.clone().requires_grad_(True)
        h = fn(a, b)
        h.backward(gy)
        outs.append((h.detach(), a.grad, b.grad))
    h64 = O.scan_fwd(gates.double().cpu().numpy(), tokens.double().cpu().numpy())
    da64, db64, _ = O.scan_bwd(gates.double().cpu().numpy(), h64, gy.double().cpu().numpy())
    for ours, theirs, ref in zip(outs[0], outs[1], (h64, da64, db64)):
        assert rel_err(ours, theirs.double().cpu().numpy()) <= TOL
        assert rel_err(ours, ref) <= TOL
        assert rel_err(theirs, ref) <= TOL       # the reference kernel itself meets the bar it sets


@pytest.mark.parametrize("T", [5, 50, 64, 200])
def test_gated_recurrent_layer_matches_reference_on_gpu(T):
    """Our GatedRecurrentLayer (h0 instead of the left pad, fused conv / gate+scan / z-gate kernels) vs the reference's
    (literal pad + Triton scan) with the same weights on the same input: output and every gradient.

    Ground truth = the reference module itself in float64 with a sequential scan.  Measured on B200 (profiles/r2_parity.md):
    this repo's fp32 path is within ~1e-6..4e-6 of it everywhere, while the REFERENCE's own fp32 GPU path (cuDNN depthwise
    conv + ~12 separate elementwise kernels + Triton scan) is 1e-4..3e-3 away in the gradients.  So the bars are: ours vs
    float64 <= 1e-4 (north star), and ours vs reference <= the reference's own distance from float64 (x2) + 1e-4."""
    import copy
    from datamining_recblr_b200.recblr import GatedRecurrentLayer
    from oracle.reference_loader import sequential_scan
    mod, ps = build_ref.load_gpu_reference()
    torch.manual_seed(T)
    ref = mod.GatedRecurrentLayer(d_model=64, expansion_factor=2, kernel_size=4).cuda()
    with torch.no_grad():
        for p in ref.parameters():
            if p.dim() >= 2:
                p.mul_(3.0)
        ref.gates.bias.normal_(std=0.5)
        ref.conv1d.bias.normal_(std=0.5)
    ours = GatedRecurrentLayer(d_model=64, expansion_factor=2, kernel_size=4).cuda()
    ours.load_state_dict(ref.state_dict())
    r64 = copy.deepcopy(ref).double()
    x = torch.randn(7, T, 64, device="cuda")
    gy = torch.randn(7, T, 64, device="cuda")
    res = {}
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False     # the fp32 contract: no TF32 in the reference's cuDNN conv fallback
    try:
        for name, m, dt in (("ours", ours, torch.float32), ("ref", ref, torch.float32), ("f64", r64, torch.float64)):
            mod.parallel_scan = sequential_scan if name == "f64" else ps.parallel_scan
            xi = x.to(dt).clone().requires_grad_(True)
            y = m(xi)
            y.backward(gy.to(dt))
            res[name] = {"y": y.detach().double(), "dx": xi.grad.double(),
                         **{n: p.grad.double().clone() for n, p in m.named_parameters()}}
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
        mod.parallel_scan = ps.parallel_scan
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
    report = {}
    for key, truth in res["f64"].items():
        e_ours, e_ref, e_pair = rel(res["ours"][key], truth), rel(res["ref"][key], truth), rel(res["ours"][key], res["ref"][key])
        report[key] = (e_ours, e_ref, e_pair)
        assert e_ours <= TOL, (key, report)
        assert e_pair <= 2 * e_ref + TOL, (key, report)
    print("layer parity T=%d (ours vs f64, reference vs f64, ours vs reference):" % T,
          {k: tuple(float(f"{x:.2g}") for x in v) for k, v in report.items()})


def test_model_loss_and_scores_match_reference_on_gpu():
    """Whole model, reference checkpoint loaded unchanged: CE loss (fused tcgen05 CE, bf16 operands: 2e-3; dense fp32:
    1e-4), seq_output and full_sort_predict scores vs the reference model running its Triton scan on the same GPU."""
    from datamining_recblr_b200.recblr import RecBLR
    from oracle import torch_port as TP
    from oracle.reference_loader import FakeDataset, make_config
    mod, _ = build_ref.load_gpu_reference()
    n_items, L, B = 700, 50, 48
    cfg = make_config(hidden_size=64, num_layers=2, dropout_prob=0.0, max_len=L)
    cfg["device"] = "cuda"
    torch.manual_seed(1)
    ref = mod.RecBLR(cfg, FakeDataset(n_items)).cuda().eval()
    ours = RecBLR(cfg, FakeDataset(n_items)).cuda().eval()
    ours.load_state_dict(ref.state_dict())
    seq, lens, pos = TP.synthetic_batch(B, L, n_items, seed=4)
    inter = {"item_id_list": seq.cuda(), "item_length": lens.cuda(), "item_id": pos.cuda()}
    torch.backends.cudnn.allow_tf32 = False   # see test_gated_recurrent_layer_matches_reference_on_gpu
    with torch.no_grad():
        q_ref = ref.forward(inter["item_id_list"], inter["item_length"])
        q = ours.forward(inter["item_id_list"], inter["item_length"])
        assert rel_err(q, q_ref.double().cpu().numpy()) <= TOL
        s_ref, s = ref.full_sort_predict(inter), ours.full_sort_predict(inter)
        assert rel_err(s, s_ref.double().cpu().numpy()) <= TOL
        l_ref = float(ref.calculate_loss(inter))
        ours.ce_impl = "dense"
        assert abs(float(ours.calculate_loss(inter)) - l_ref) <= TOL * abs(l_ref)
        ours.ce_impl = "fused"
        assert abs(float(ours.calculate_loss(inter)) - l_ref) <= 2e-3 * abs(l_ref)
        # ranking: (a) our fp32 dense scores rank like the reference's (fp32 near-ties aside); (b) the fused scorer fed the
        # REFERENCE's seq_output returns exactly the ids of RecBole's dense procedure on the same bf16-rounded operands
        from datamining_recblr_b200 import ops
        top = lambda m: torch.sort(m, dim=1, descending=True, stable=True).indices[:, :10]
        sd, so = s_ref.double().clone(), s.double().clone()
        sd[:, 0] = so[:, 0] = float("-inf")
        assert float((top(sd) == top(so)).all(1).double().mean()) >= 0.95
        table = ref.item_embedding.weight.detach()
        _, ids = ops.fullsort_topk(q_ref, table, 10, mask_id=0)
        dense = q_ref.bfloat16().double() @ table.bfloat16().double().T
        dense[:, 0] = float("-inf")
        assert torch.equal(ids.long(), top(dense))


def test_level1_adoption_unmodified_reference_module_over_the_two_shims():
    """INTEGRATION.md level 1: the UNMODIFIED RecBLR.py with only its two imported kernels swapped for this repo's drop-ins
    (`parallel_scan` -> csrc/scan_bct.cu, `causal_conv1d_fn` -> csrc/conv1d.cu) — loss and every parameter gradient equal
    the same module running its own Triton scan + F.conv1d fallback."""
    from datamining_recblr_b200.causal_conv1d import causal_conv1d_fn
    from datamining_recblr_b200.parallel_scan import parallel_scan
    from oracle import torch_port as TP
    from oracle.reference_loader import FakeDataset, make_config
    mod, ps = build_ref.load_gpu_reference()
    n_items, L, B = 400, 50, 24
    cfg = make_config(hidden_size=64, num_layers=2, dropout_prob=0.0, max_len=L)
    cfg["device"] = "cuda"
    torch.manual_seed(2)
    model = mod.RecBLR(cfg, FakeDataset(n_items)).cuda().train()
    seq, lens, pos = TP.synthetic_batch(B, L, n_items, seed=6)
    inter = {"item_id_list": seq.cuda(), "item_length": lens.cuda(), "item_id": pos.cuda()}
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    res = []
    try:
        for scan, conv in ((ps.parallel_scan, None), (parallel_scan, causal_conv1d_fn)):
            mod.parallel_scan, mod.causal_conv1d_fn = scan, conv
            model.zero_grad(set_to_none=True)
            loss = model.calculate_loss(inter)
            loss.backward()
            res.append((float(loss), {n: p.grad.clone() for n, p in model.named_parameters()}))
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
        mod.parallel_scan, mod.causal_conv1d_fn = ps.parallel_scan, None
    assert abs(res[0][0] - res[1][0]) <= 1e-5 * abs(res[0][0])
    for n, g in res[0][1].items():
        den = float(g.abs().max()) + 1e-12
        assert float((g - res[1][1][n]).abs().max()) <= 2e-3 * den, n     # the reference's own fp32 GPU path is ~1e-3 noisy
