import numpy as np
import torch


def rel_err(x, ref):
    """max |x - ref| / max |ref|  — the 'max relative error' of BASELINE.json's north_star, normalised by the
    largest reference magnitude so that near-zero entries do not dominate."""
    x = x.detach().double().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert x.shape == ref.shape, (x.shape, ref.shape)
    den = max(np.abs(ref).max(), 1e-30)
    return float(np.abs(x - ref).max() / den)


def cuda(a, dtype=torch.float32, requires_grad=False):
    t = torch.as_tensor(np.asarray(a), dtype=dtype, device="cuda")
    return t.requires_grad_(requires_grad)


def elem_rel_err(x, ref, floor=1e-3):
    """Element-wise relative error  max_j |x_j - ref_j| / max(|ref_j|, floor * max|ref|):  the literal reading of the
    north star's "max relative error", with an absolute floor (a fraction of the largest reference magnitude) so that
    entries that are zero up to rounding do not divide by ~0.  Stricter than rel_err for small-magnitude outputs."""
    x = x.detach().double().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref.detach().double().cpu().numpy() if isinstance(ref, torch.Tensor) else ref, dtype=np.float64)
    assert x.shape == ref.shape, (x.shape, ref.shape)
    den = np.maximum(np.abs(ref), floor * max(np.abs(ref).max(), 1e-30))
    return float((np.abs(x - ref) / den).max())
