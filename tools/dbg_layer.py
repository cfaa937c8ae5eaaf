"""debug: ours vs reference GatedRecurrentLayer vs float64 on GPU"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import build_ref
from datamining_recblr_b200.recblr import GatedRecurrentLayer
mod, ps = build_ref.load_gpu_reference()
torch.backends.cudnn.allow_tf32 = False
for T in (64, 50):
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
    # float64 copy of the reference with a sequential scan
    import copy
    r64 = copy.deepcopy(ref).double()
    from oracle.reference_loader import sequential_scan
    x = torch.randn(7, T, 64, device="cuda")
    gy = torch.randn(7, T, 64, device="cuda")
    outs = {}
    for name, m, dt in (("ours", ours, torch.float32), ("ref", ref, torch.float32), ("f64", r64, torch.float64)):
        if name == "f64":
            mod.parallel_scan = sequential_scan
        else:
            mod.parallel_scan = ps.parallel_scan
        xi = x.to(dt).clone().requires_grad_(True)
        y = m(xi)
        y.backward(gy.to(dt))
        outs[name] = (y.detach().double(), xi.grad.double(), {n: p.grad.double().clone() for n, p in m.named_parameters()})
    def rel(a, b):
        return float((a - b).abs().max() / b.abs().max())
    print("T", T, "y: ours", rel(outs["ours"][0], outs["f64"][0]), "ref", rel(outs["ref"][0], outs["f64"][0]))
    print("   dx: ours", rel(outs["ours"][1], outs["f64"][1]), "ref", rel(outs["ref"][1], outs["f64"][1]))
    for n in outs["f64"][2]:
        print("   ", n, "ours", rel(outs["ours"][2][n], outs["f64"][2][n]), "ref", rel(outs["ref"][2][n], outs["f64"][2][n]))
