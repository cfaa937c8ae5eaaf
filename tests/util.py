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
