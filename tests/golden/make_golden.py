"""Generates tests/golden/*.npz from the UNMODIFIED reference (/root/reference/RecBLR.py), run in the
build container through oracle/reference_loader.py (RecBole stub, F.conv1d fallback, sequential scan).
The reference ships no tests or golden vectors, so these reference-generated fixtures are
what pins the oracle.  Re-run:  python tests/golden/make_golden.py     (needs /root/reference)

All tensors float64 (model.double()) so the fixtures are exact to ~1e-15; eval mode / dropout 0.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_loader as rl  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def npy(t):
    return t.detach().cpu().numpy()


def grad_or_zero(p):
    return npy(p.grad) if p.grad is not None else np.zeros(tuple(p.shape))


def layer_case(ref, T, B=2, d_model=16, seed=0, disable_conv1d=False):
    torch.manual_seed(seed)
    layer = ref.GatedRecurrentLayer(d_model=d_model, expansion_factor=2, kernel_size=4,
                                    disable_conv1d=disable_conv1d).double()
    # non-trivial gate bias / Lambda so every gradient path is exercised
    with torch.no_grad():
        layer.gates.bias.normal_(std=0.5)
        layer.gates.weight.normal_(std=0.3)
        layer.input.weight.normal_(std=0.5)
        layer.output.weight.normal_(std=0.5)
    x = torch.randn(B, T, d_model, dtype=torch.float64, requires_grad=True)
    y = layer(x)
    gy = torch.randn_like(y)
    y.backward(gy)
    d = {"x": npy(x), "y": npy(y), "gy": npy(gy), "dx": npy(x.grad)}
    for k, v in layer.state_dict().items():
        d["p." + k] = npy(v)
    for k, v in layer.named_parameters():
        d["g." + k] = grad_or_zero(v)
    return d


def model_case(ref, n_items=50, L=12, B=6, hidden=16, seed=1, **flags):
    torch.manual_seed(seed)
    cfg = rl.make_config(hidden_size=hidden, num_layers=2, dropout_prob=0.0, max_len=L, **flags)
    model = ref.RecBLR(cfg, rl.FakeDataset(n_items)).double()
    with torch.no_grad():  # the default N(0, 0.02) init makes everything nearly linear; widen it
        for p in model.parameters():
            if p.dim() >= 2:
                p.mul_(10.0)
        for lyr in model.recurrent_layers:
            lyr.behavior_modeling.gates.bias.normal_(std=0.5)
    model.eval()
    lens = torch.randint(1, L + 1, (B,))
    lens[0] = L
    lens[1] = 1
    seq = torch.zeros(B, L, dtype=torch.long)
    for b in range(B):
        seq[b, :lens[b]] = torch.randint(1, n_items, (int(lens[b]),))
    pos = torch.randint(1, n_items, (B,))
    inter = {"item_id_list": seq, "item_length": lens, "item_id": pos}
    seq_out = model.forward(seq, lens)
    scores = model.full_sort_predict(inter)
    loss = model.calculate_loss(inter)
    model.zero_grad()
    loss.backward()
    d = {"item_seq": npy(seq), "item_len": npy(lens), "pos": npy(pos), "seq_output": npy(seq_out),
         "scores": npy(scores), "loss": npy(loss), "n_layers": np.int64(2)}
    for k, v in model.state_dict().items():
        d["p." + k] = npy(v)
    for k, v in model.named_parameters():
        d["g." + k] = grad_or_zero(v)
    return d


def main():
    ref = rl.load_reference_module()
    for T in (1, 5, 50, 200):
        np.savez_compressed(os.path.join(OUT, f"ref_layer_T{T}.npz"), **layer_case(ref, T, seed=T))
    np.savez_compressed(os.path.join(OUT, "ref_layer_T50_noconv.npz"),
                        **layer_case(ref, 50, seed=7, disable_conv1d=True))
    np.savez_compressed(os.path.join(OUT, "ref_model_small.npz"), **model_case(ref))
    np.savez_compressed(os.path.join(OUT, "ref_model_small_bdlru_only.npz"),
                        **model_case(ref, seed=3, bd_lru_only=True))
    # parameter count of the logged configuration (log/RecBLR/...12-51-02...log:143)
    cfg = rl.make_config(hidden_size=64, num_layers=2, dropout_prob=0.2, max_len=200)
    m = ref.RecBLR(cfg, rl.FakeDataset(10544))
    n = sum(p.numel() for p in m.parameters() if p.requires_grad)
    shapes = {k: np.array(v.shape, dtype=np.int64) for k, v in m.state_dict().items()}
    np.savez_compressed(os.path.join(OUT, "ref_state_shapes.npz"), n_params=np.int64(n), **shapes)
    print("trainable parameters:", n)


if __name__ == "__main__":
    main()
