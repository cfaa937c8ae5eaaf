"""CPU pins of the item-bias scorer: the oracle's biased logits / CE / gradients against the literal torch
expression of bert4rec.py:200-213, 230-242 in float64, and the exact bf16 split the CUDA path uses to carry the bias."""
import numpy as np
import torch

from oracle import bdlru_oracle as O


def test_oracle_biased_ce_matches_reference_expression():
    rng = np.random.default_rng(3)
    B, N, D = 17, 41, 8
    q = torch.tensor(rng.normal(size=(B, D)), requires_grad=True)
    e = torch.tensor(rng.normal(size=(N, D)), requires_grad=True)
    b = torch.tensor(rng.normal(size=N), requires_grad=True)
    pos = rng.integers(0, N, size=B)
    # bert4rec.py:200-213 with every target weight = 1 (the caller selects the masked positions)
    logits = torch.matmul(q, e.transpose(0, 1)) + b
    loss_fct = torch.nn.CrossEntropyLoss(reduction="none")
    targets = torch.ones(B, dtype=torch.float64)
    loss = torch.sum(loss_fct(logits, torch.tensor(pos)) * targets) / torch.sum(targets)
    loss.backward()
    l, lse, dQ, dE = O.ce_loss(q.detach().numpy(), e.detach().numpy(), pos, b.detach().numpy())
    db = O.ce_bias_grad(q.detach().numpy(), e.detach().numpy(), pos, b.detach().numpy())
    assert abs(l - float(loss)) < 1e-12
    assert np.abs(dQ - q.grad.numpy()).max() < 1e-12 and np.abs(dE - e.grad.numpy()).max() < 1e-12
    assert np.abs(db - b.grad.numpy()).max() < 1e-12
    s = O.full_sort_scores(q.detach().numpy(), e.detach().numpy(), b.detach().numpy())
    assert np.abs(s - logits.detach().numpy()).max() < 1e-12


def test_bias_split_is_exact_and_augmented_gemm_equals_biased_scores():
    from datamining_recblr_b200 import ops
    torch.manual_seed(0)
    N, D = 5000, 64
    bias = torch.randn(N) * 3
    bias[:4] = torch.tensor([0.0, 1e-30, -65504.0, 3.0e38])
    qb, eb = torch.randn(6, D).bfloat16(), torch.randn(N, D).bfloat16()
    qa, ea = ops._augment_with_bias(qb, eb, bias)
    assert qa.shape == (6, D + ops.BIAS_COLS) and ea.shape == (N, D + ops.BIAS_COLS)
    rec = (ea[:, D].double() + ea[:, D + 1].double()) + ea[:, D + 2].double()
    assert (rec == bias.double()).all()
    assert (qa[:, D:D + 3] == 1).all() and (qa[:, D + 3:] == 0).all() and (ea[:, D + 3:] == 0).all()
    ok = torch.isfinite(bias) & (bias.abs() < 1e30)
    s0 = qb.double() @ eb.double().T + bias.double()
    s1 = qa.double() @ ea.double().T
    assert (s0[:, ok] == s1[:, ok]).all()


def test_baseline_heads_fail_loudly_on_cpu_tensors():
    """scoring.py is CUDA-only: no dense fallback may answer for CPU tensors (the product path must not degrade silently)."""
    import pytest
    from datamining_recblr_b200 import scoring
    from datamining_recblr_b200._lib import BdlruError
    q, w, b = torch.randn(4, 64), torch.randn(30, 64), torch.randn(30)
    pos = torch.randint(0, 30, (4,))
    with pytest.raises(BdlruError):
        scoring.cross_entropy(q, w, pos, output_bias=b)
    with pytest.raises(BdlruError):
        scoring.full_sort_topk(q, w, 5, output_bias=b)
