"""TEST INFRASTRUCTURE ONLY — loads the UNMODIFIED reference model from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  Used by
tests/golden/make_golden.py to generate the committed fixtures and by the CPU tests (when the
reference tree is present) to pin oracle/bdlru_oracle.py against the reference itself.

What is substituted, and why:
  * `recbole` is not installed -> oracle/recbole_stub provides the two imports of RecBLR.py:4-5.
  * `causal_conv1d` is not installed -> RecBLR.py:8-11 falls back to its own F.conv1d line (185).
  * The Triton kernels of parallel_scan.py cannot launch without a GPU -> after import,
    `RecBLR.parallel_scan` is re-bound to `sequential_scan` below, a plain autograd loop with the
    semantics of parallel_scan.py:35-41 (h_t = a_t*h_{t-1} + x_t, h_{-1} = 0) on [B, C, T].
"""
import importlib
import os
import sys

import torch

REFERENCE_ROOT = os.environ.get("RECBLR_REFERENCE_ROOT", "/root/reference")
_STUB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "recbole_stub")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "RecBLR.py"))


def sequential_scan(gates: torch.Tensor, tokens: torch.Tensor) -> torch.Tensor:
    """h_t = gates_t * h_{t-1} + tokens_t over the last axis of [B, C, T] (parallel_scan.py:35-41)."""
    assert gates.shape == tokens.shape
    h = torch.zeros_like(tokens[..., 0])
    out = []
    for t in range(tokens.shape[-1]):
        h = gates[..., t] * h + tokens[..., t]
        out.append(h)
    return torch.stack(out, dim=-1)


def load_reference_module():
    """Import /root/reference/RecBLR.py as module `RecBLR` with the scan re-bound to the CPU loop."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (_STUB, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    mod = importlib.import_module("RecBLR")
    mod.parallel_scan = sequential_scan
    return mod


class FakeDataset:
    """`dataset.num(ITEM_ID)` is the only dataset call the model makes."""

    def __init__(self, n_items):
        self.n_items = n_items

    def num(self, field):
        return self.n_items


class FakeConfig(dict):
    """RecBole's Config returns None for unknown keys."""

    def __getitem__(self, k):
        return self.get(k, None)


def make_config(hidden_size=64, num_layers=2, dropout_prob=0.2, expand=2, d_conv=4, loss_type="CE",
                max_len=200, **flags):
    cfg = FakeConfig(
        hidden_size=hidden_size, num_layers=num_layers, dropout_prob=dropout_prob, expand=expand,
        d_conv=d_conv, loss_type=loss_type, USER_ID_FIELD="user_id", ITEM_ID_FIELD="item_id",
        LIST_SUFFIX="_list", ITEM_LIST_LENGTH_FIELD="item_length", NEG_PREFIX="neg_",
        MAX_ITEM_LIST_LENGTH=max_len, device="cpu")
    cfg.update(flags)
    return cfg
