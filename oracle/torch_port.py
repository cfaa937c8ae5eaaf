"""TEST INFRASTRUCTURE ONLY — plain-PyTorch CPU restatement of the reference's train / eval step, with
autograd, used (a) as the CPU baseline that bench.py times on the GPU box's host cores (`cpu_baseline`,
`--impl reference`; /root/reference does not exist there) and (b) by tests as a second checker.  The product
package never imports it.

It is "the reference's PyTorch path with the sequential scan on CPU" of BASELINE.json configs[0]: the
literal left zero-pad to a power of two (RecBLR.py:177-179), the F.conv1d fallback line (RecBLR.py:185), the
separate elementwise gate ops (197-199), both transposes (200) and a sequential loop with the semantics of
parallel_scan.py:35-41 in place of the Triton kernel (which cannot run on a CPU).  Functional style: the
weights come in as a dict keyed by the reference's state_dict names.

Pinned by tests/test_oracle_golden.py against the reference-generated fixtures in tests/golden/.
"""
import torch
import torch.nn.functional as F


def seq_scan(gates, tokens):
    """h_t = gates_t * h_{t-1} + tokens_t along the last axis of [B, C, T], h_{-1} = 0 (parallel_scan.py:35-41)."""
    h = torch.zeros_like(tokens[..., 0])
    out = []
    for t in range(tokens.shape[-1]):
        h = gates[..., t] * h + tokens[..., t]
        out.append(h)
    return torch.stack(out, dim=-1)


def bdlru_layer(x, w, pre, disable_conv1d=False):
    """GatedRecurrentLayer.forward, RecBLR.py:170-207, literal padding."""
    T = x.shape[1]
    xz = F.linear(x, w[pre + "input.weight"])                       # :173
    xs, z = xz.chunk(2, dim=-1)                                     # :174
    P = 2 ** ((T - 1).bit_length()) - T                             # :177
    if P:
        xs = F.pad(xs, (0, 0, P, 0))                                # :179
    if not disable_conv1d:                                          # :185
        cw = w[pre + "conv1d.weight"]
        y = F.conv1d(xs.mT, cw, w[pre + "conv1d.bias"], padding=cw.shape[-1] - 1, groups=cw.shape[0])
        xs = F.silu(y[..., :T + P].mT)
    rec, inp = F.linear(xs, w[pre + "gates.weight"], w[pre + "gates.bias"]).chunk(2, dim=-1)   # :196
    alpha = torch.exp(-F.softplus(w[pre + "Lambda"]) * torch.sigmoid(rec))                    # :197
    beta = torch.sqrt(1 - alpha.pow(2) + 1e-8) * torch.sigmoid(inp)                           # :198
    h = seq_scan(alpha.mT.contiguous(), (beta * xs).mT.contiguous()).mT                       # :199-200
    if P:
        h = h[:, P:]                                                # :204
    return F.linear(F.silu(z) * h, w[pre + "output.weight"])        # :206


def _ln(x, w, pre):
    return F.layer_norm(x, x.shape[-1:], w[pre + "weight"], w[pre + "bias"], eps=1e-12)


def seq_output(w, item_seq, item_len, num_layers, disable_conv1d=False, disable_ffn=False, dropout_p=0.0):
    """RecBLR.forward (RecBLR.py:75-84) with RecurrentLayer.forward (140-145) and FeedForward.forward (218-227)."""
    drop = (lambda t: F.dropout(t, dropout_p, True)) if dropout_p > 0 else (lambda t: t)
    x = _ln(drop(w["item_embedding.weight"][item_seq]), w, "layer_norm.")
    for l in range(num_layers):
        pre = f"recurrent_layers.{l}."
        y = bdlru_layer(x, w, pre + "behavior_modeling.", disable_conv1d)
        x = _ln(drop(y) + x, w, pre + "layer_norm.")
        if not disable_ffn:
            h = drop(F.silu(F.linear(x, w[pre + "ffn.w_1.weight"], w[pre + "ffn.w_1.bias"])))
            h = drop(F.linear(h, w[pre + "ffn.w_2.weight"], w[pre + "ffn.w_2.bias"]))
            x = _ln(h + x, w, pre + "ffn.layer_norm.")
    idx = (item_len - 1).view(-1, 1, 1).expand(-1, 1, x.shape[-1])
    return x.gather(1, idx).squeeze(1)


def ce_loss(w, item_seq, item_len, pos, num_layers, **kw):
    """calculate_loss, CE branch (RecBLR.py:99-103): dense logits over all rows + mean cross-entropy."""
    q = seq_output(w, item_seq, item_len, num_layers, **kw)
    return F.cross_entropy(q @ w["item_embedding.weight"].T, pos)


def full_sort_topk(w, item_seq, item_len, num_layers, k, **kw):
    """full_sort_predict (RecBLR.py:114-122) + RecBole's scores[:, 0] = -inf and top-k ([upstream]),
    ties broken by lowest index (stable sort)."""
    q = seq_output(w, item_seq, item_len, num_layers, **kw)
    s = q @ w["item_embedding.weight"].T
    s[:, 0] = float("-inf")
    order = torch.sort(s, dim=1, descending=True, stable=True).indices[:, :k]
    return s.gather(1, order), order


def init_weights(n_items, hidden, num_layers, expand=2, d_conv=4, seed=2020, dtype=torch.float32):
    """Random weights with the reference's shapes and init distributions (RecBLR.py:66-73, 153-167; Conv1d keeps
    PyTorch's default U(-1/sqrt(k), 1/sqrt(k)) init).  For synthetic benchmarks and tests."""
    g = torch.Generator().manual_seed(seed)
    C, D = hidden * expand, hidden
    n = lambda *s: torch.randn(*s, generator=g, dtype=dtype) * 0.02
    w = {"item_embedding.weight": n(n_items, D), "layer_norm.weight": torch.ones(D, dtype=dtype),
         "layer_norm.bias": torch.zeros(D, dtype=dtype)}
    inv = lambda y: torch.log(torch.expm1(torch.tensor(y, dtype=torch.float64)))
    lo, hi = inv(-torch.log(torch.tensor(0.9)).item()).item(), inv(-torch.log(torch.tensor(0.999)).item()).item()
    bound = 1.0 / (d_conv ** 0.5)
    for l in range(num_layers):
        p = f"recurrent_layers.{l}."
        b = p + "behavior_modeling."
        w[b + "Lambda"] = torch.linspace(lo, hi, C, dtype=dtype)
        w[b + "input.weight"] = n(2 * C, D)
        w[b + "conv1d.weight"] = (torch.rand(C, 1, d_conv, generator=g, dtype=dtype) * 2 - 1) * bound
        w[b + "conv1d.bias"] = (torch.rand(C, generator=g, dtype=dtype) * 2 - 1) * bound
        w[b + "gates.weight"] = n(2 * C, C)
        w[b + "gates.bias"] = torch.zeros(2 * C, dtype=dtype)
        w[b + "output.weight"] = n(D, C)
        w[p + "layer_norm.weight"], w[p + "layer_norm.bias"] = torch.ones(D, dtype=dtype), torch.zeros(D, dtype=dtype)
        w[p + "ffn.w_1.weight"], w[p + "ffn.w_1.bias"] = n(4 * D, D), torch.zeros(4 * D, dtype=dtype)
        w[p + "ffn.w_2.weight"], w[p + "ffn.w_2.bias"] = n(D, 4 * D), torch.zeros(D, dtype=dtype)
        w[p + "ffn.layer_norm.weight"] = torch.ones(D, dtype=dtype)
        w[p + "ffn.layer_norm.bias"] = torch.zeros(D, dtype=dtype)
    return w


def synthetic_batch(B, L, n_items, seed=2020):
    """ids uniform in [1, n_items), lengths uniform in [min(5, L), L], right-padded with 0 (RecBole's convention),
    targets uniform in [1, n_items)  (SURVEY §8d 'Training inputs')."""
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(min(5, L), L + 1, (B,), generator=g)
    seq = torch.randint(1, n_items, (B, L), generator=g)
    seq = seq * (torch.arange(L)[None, :] < lens[:, None])
    pos = torch.randint(1, n_items, (B,), generator=g)
    return seq, lens, pos
