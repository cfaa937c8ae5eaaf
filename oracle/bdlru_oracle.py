"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy, float64 by default) of the reference's
BD-LRU hot path.  It is the checker for the CUDA path, never the thing shipped or measured:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import it.  The product package (datamining_recblr_b200/) never does.

Every function cites the reference lines it restates (paths relative to /root/reference).

Parity pin: the reference holds no tests or golden vectors for this path,
so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the build container
through oracle/reference_loader.py (unmodified RecBLR.py + RecBole stub + sequential scan);
the resulting fixtures are committed under tests/golden/ with their generator
(tests/golden/make_golden.py) and checked by tests/test_oracle_golden.py.
"""
import numpy as np

F64 = np.float64


# ----------------------------------------------------------------------------- elementwise
def sigmoid(x):
    x = np.asarray(x)
    return 1.0 / (1.0 + np.exp(-x))


def softplus(x):
    x = np.asarray(x)
    return np.logaddexp(0.0, x)


def silu(x):
    return x * sigmoid(x)


def silu_grad(x):
    s = sigmoid(x)
    return s * (1.0 + x * (1.0 - s))


def lambda_init(C, r_min=0.9, r_max=0.999):
    """Lambda = linspace(softplus^-1(-ln r_min), softplus^-1(-ln r_max), C)  (RecBLR.py:153-158,166)."""
    inv = lambda y: np.log(np.expm1(y))
    return np.linspace(inv(-np.log(r_min)), inv(-np.log(r_max)), C).astype(F64)


# ----------------------------------------------------------------------------- raw scan (S0)
def scan_fwd(gates, tokens, h0=None):
    """Inclusive first-order scan over the LAST axis: h_t = gates_t*h_{t-1} + tokens_t,
    h_{-1} = h0 (0 in the reference).  parallel_scan.py:35-41 (operator), 85-95 (forward)."""
    gates = np.asarray(gates)
    tokens = np.asarray(tokens)
    assert gates.shape == tokens.shape
    h = np.zeros(tokens.shape[:-1], tokens.dtype) if h0 is None else np.asarray(h0, tokens.dtype).copy()
    out = np.empty_like(tokens)
    for t in range(tokens.shape[-1]):
        h = gates[..., t] * h + tokens[..., t]
        out[..., t] = h
    return out


def scan_bwd(gates, states, grad_out, h0=None):
    """Backward identities of parallel_scan.py:98-114:
        dh~_t = g_t + a_{t+1} * dh~_{t+1}   (reverse scan with gates shifted by one, last = 1)
        d_tokens = dh~ ;  d_gates_t = h_{t-1} * dh~_t  (h_{-1} = h0 = 0 in the reference).
    Returns (d_gates, d_tokens, d_h0)."""
    gates = np.asarray(gates)
    T = gates.shape[-1]
    d = np.zeros(gates.shape[:-1], gates.dtype)
    d_states = np.empty_like(gates)
    for t in range(T - 1, -1, -1):
        nxt = gates[..., t + 1] * d if t + 1 < T else 0.0
        d = grad_out[..., t] + nxt
        d_states[..., t] = d
    prev = np.zeros_like(states)
    prev[..., 1:] = states[..., :-1]
    if h0 is not None:
        prev[..., 0] = h0
    d_gates = prev * d_states
    d_h0 = gates[..., 0] * d_states[..., 0]
    return d_gates, d_states, d_h0


# ----------------------------------------------------------------------------- gate math + fused scan (S1)
def gate_math(xp, r, i, Lambda):
    """alpha = exp(-softplus(Lambda)*sigmoid(r)); beta = sqrt(1-alpha^2+1e-8)*sigmoid(i);
    beta_prime = beta*x'.  RecBLR.py:197-199.  Channel-last [..., C]."""
    alpha = np.exp(-softplus(Lambda) * sigmoid(r))
    beta = np.sqrt(1.0 - alpha ** 2 + 1e-8) * sigmoid(i)
    return alpha, beta * xp


def gated_scan_fwd(xp, r, i, Lambda, h0=None):
    """Fused S1 boundary on channel-last [B, T, C]: gate math (RecBLR.py:197-199) then the scan
    of RecBLR.py:200 (here along axis 1, no transposes), started from h0[C] or h0[B,C]."""
    a, bp = gate_math(xp, r, i, Lambda)
    h0b = None
    if h0 is not None:
        h0b = np.broadcast_to(np.asarray(h0, a.dtype), (a.shape[0], a.shape[2])).copy()
    h = scan_fwd(np.swapaxes(a, 1, 2), np.swapaxes(bp, 1, 2), h0b)
    return np.swapaxes(h, 1, 2)


def gated_scan_bwd(xp, r, i, Lambda, h0, grad_h):
    """Analytic backward of gated_scan_fwd (chain rule through RecBLR.py:197-199 and the
    identities of parallel_scan.py:98-114).  Returns (dxp, dr, di, dLambda, dh0[C])."""
    B, T, C = xp.shape
    sr, si, c = sigmoid(r), sigmoid(i), softplus(Lambda)
    a = np.exp(-c * sr)
    q = np.sqrt(1.0 - a ** 2 + 1e-8)
    beta = q * si
    h0b = np.zeros((B, C), xp.dtype) if h0 is None else np.broadcast_to(np.asarray(h0, xp.dtype), (B, C))
    h = gated_scan_fwd(xp, r, i, Lambda, h0)
    da, dbp, dh0 = scan_bwd(np.swapaxes(a, 1, 2), np.swapaxes(h, 1, 2), np.swapaxes(grad_h, 1, 2), h0b)
    da, dbp = np.swapaxes(da, 1, 2), np.swapaxes(dbp, 1, 2)
    dxp = dbp * beta
    dbeta = dbp * xp
    di = dbeta * q * si * (1.0 - si)
    da_tot = da - dbeta * si * a / q
    dsr = da_tot * a * (-c)
    dr = dsr * sr * (1.0 - sr)
    dc = (da_tot * a * (-sr)).sum(axis=(0, 1))
    dLambda = dc * sigmoid(Lambda)
    return dxp, dr, di, dLambda, dh0.sum(axis=0)


# ----------------------------------------------------------------------------- causal depthwise conv + SiLU
def causal_conv1d_silu_fwd(x, weight, bias, activation=True):
    """y_t = silu(bias + sum_j w[:, j] * x_{t-(W-1)+j}) on channel-last [B, T, C], zeros before t=0.
    Restates the fallback line RecBLR.py:185 (conv1d with padding=W-1 truncated to T) which pins the
    semantics of causal_conv1d_fn (RecBLR.py:188-193).  weight is [C, W]."""
    B, T, C = x.shape
    W = weight.shape[1]
    xp = np.concatenate([np.zeros((B, W - 1, C), x.dtype), x], axis=1)
    pre = np.zeros_like(x) + (0.0 if bias is None else bias)
    for j in range(W):
        pre = pre + xp[:, j:j + T, :] * weight[:, j]
    return silu(pre) if activation else pre


def causal_conv1d_silu_bwd(x, weight, bias, grad_y, activation=True):
    """Returns (dx, dweight[C,W], dbias[C])."""
    B, T, C = x.shape
    W = weight.shape[1]
    pre = causal_conv1d_silu_fwd(x, weight, bias, activation=False)
    dpre = grad_y * silu_grad(pre) if activation else grad_y
    xp = np.concatenate([np.zeros((B, W - 1, C), x.dtype), x], axis=1)
    dxp = np.zeros_like(xp)
    dw = np.zeros_like(weight)
    for j in range(W):
        dxp[:, j:j + T, :] += dpre * weight[:, j]
        dw[:, j] = (dpre * xp[:, j:j + T, :]).sum(axis=(0, 1))
    return dxp[:, W - 1:, :], dw, dpre.sum(axis=(0, 1))


# ----------------------------------------------------------------------------- left-pad quirk
def left_pad_len(T):
    """pad_len = 2**ceil(log2 T) - T  (RecBLR.py:177)."""
    return 2 ** ((T - 1).bit_length()) - T


def phantom_h0(P, conv_bias, gates_w, gates_b, Lambda, disable_conv1d=False):
    """State entering the first real step after the reference's P left-padded phantom steps
    (RecBLR.py:177-199): on padded steps the conv output is silu(conv_bias), so
    h0 = b' * (1 - a^P) / (1 - a) with a, b' evaluated at that constant vector."""
    C = Lambda.shape[0]
    if P == 0 or disable_conv1d:
        return np.zeros(C, Lambda.dtype)
    s = silu(conv_bias)
    g = gates_w @ s + gates_b
    a, bp = gate_math(s, g[:C], g[C:], Lambda)
    return bp * (1.0 - a ** P) / (1.0 - a)


# ----------------------------------------------------------------------------- layers (eval mode)
def layer_norm(x, w, b, eps=1e-12):
    """nn.LayerNorm(eps=1e-12) over the last axis (RecBLR.py:41,137,216)."""
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * w + b


def gated_recurrent_layer(x, p, disable_conv1d=False, literal_padding=True):
    """GatedRecurrentLayer.forward (RecBLR.py:170-207) on [B, T, D].  p: dict with the module's
    state_dict entries (input.weight, conv1d.weight [C,1,W], conv1d.bias, gates.weight, gates.bias,
    Lambda, output.weight).  literal_padding=True replays the reference's left zero-pad to a power of
    two; False uses the equivalent unpadded form with phantom_h0 (what the CUDA path computes)."""
    B, T, D = x.shape
    C = p["Lambda"].shape[0]
    xz = x @ p["input.weight"].T                                   # :173
    xs, z = xz[..., :C], xz[..., C:]                               # :174
    P = left_pad_len(T)                                            # :177
    cw = p["conv1d.weight"].reshape(C, -1)
    if literal_padding:
        xs = np.concatenate([np.zeros((B, P, C), x.dtype), xs], axis=1)   # :179
        h0 = None
    else:
        h0 = phantom_h0(P, p["conv1d.bias"], p["gates.weight"], p["gates.bias"], p["Lambda"], disable_conv1d)
    if not disable_conv1d:
        xs = causal_conv1d_silu_fwd(xs, cw, p["conv1d.bias"])      # :185
    g = xs @ p["gates.weight"].T + p["gates.bias"]                 # :196
    h = gated_scan_fwd(xs, g[..., :C], g[..., C:], p["Lambda"], h0)  # :197-200
    if literal_padding:
        h = h[:, P:]                                               # :204
    return (silu(z) * h) @ p["output.weight"].T                    # :206


def feed_forward(x, p):
    """FeedForward.forward in eval mode (RecBLR.py:218-227)."""
    h = silu(x @ p["w_1.weight"].T + p["w_1.bias"])
    h = h @ p["w_2.weight"].T + p["w_2.bias"]
    return layer_norm(h + x, p["layer_norm.weight"], p["layer_norm.bias"])


def _sub(state, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in state.items() if k.startswith(prefix)}


def recblr_forward(state, item_seq, item_seq_len, num_layers, disable_conv1d=False, disable_ffn=False,
                   literal_padding=True):
    """RecBLR.forward in eval mode (RecBLR.py:75-84; RecurrentLayer.forward 140-145).  `state` maps
    the reference state_dict names to numpy arrays.  Returns seq_output [B, D]."""
    x = state["item_embedding.weight"][item_seq]                                   # :76
    x = layer_norm(x, state["layer_norm.weight"], state["layer_norm.bias"])        # :78
    for l in range(num_layers):
        pre = f"recurrent_layers.{l}."
        y = gated_recurrent_layer(x, _sub(state, pre + "behavior_modeling."), disable_conv1d, literal_padding)
        x = layer_norm(y + x, state[pre + "layer_norm.weight"], state[pre + "layer_norm.bias"])   # :142
        if not disable_ffn:
            x = feed_forward(x, _sub(state, pre + "ffn."))                          # :144
    idx = np.asarray(item_seq_len) - 1                                             # :83
    return x[np.arange(x.shape[0]), idx]


# ----------------------------------------------------------------------------- scoring, CE, full-sort eval
def full_sort_scores(seq_output, item_emb, item_bias=None):
    """scores = seq_output @ E^T  -> [B, n_items]   (RecBLR.py:118-122); `+ output_bias` for the BERT4Rec baseline
    (bert4rec.py:230-242)."""
    s = seq_output @ item_emb.T
    return s if item_bias is None else s + item_bias[None, :]


def ce_loss(seq_output, item_emb, pos_items, item_bias=None):
    """Mean cross-entropy over ALL n_items rows incl. row 0 (RecBLR.py:99-103; with `item_bias`: bert4rec.py:200-213
    restricted to the rows whose target weight is 1).
    Returns (loss, lse[B], dQ[B,D], dE[N,D]) with gradients of the mean loss; d(item_bias) = column sums of the
    softmax-minus-onehot matrix = `ce_bias_grad`."""
    logits = full_sort_scores(seq_output, item_emb, item_bias)
    m = logits.max(axis=1, keepdims=True)
    lse = (m + np.log(np.exp(logits - m).sum(axis=1, keepdims=True)))[:, 0]
    B = logits.shape[0]
    loss = (lse - logits[np.arange(B), pos_items]).mean()
    p = np.exp(logits - lse[:, None])
    p[np.arange(B), pos_items] -= 1.0
    p /= B
    return loss, lse, p @ item_emb, p.T @ seq_output


def ce_bias_grad(seq_output, item_emb, pos_items, item_bias):
    """d(mean CE)/d(item_bias) [N] for logits = Q E^T + bias (bert4rec.py:200-213)."""
    logits = full_sort_scores(seq_output, item_emb, item_bias)
    m = logits.max(axis=1, keepdims=True)
    p = np.exp(logits - m)
    p /= p.sum(axis=1, keepdims=True)
    B = logits.shape[0]
    p[np.arange(B), pos_items] -= 1.0
    return p.sum(axis=0) / B


def topk_lowest_index(scores, k, mask_col0=True):
    """RecBole's full-sort eval step ([upstream] Trainer._full_sort_batch_eval + Collector):
    scores[:, 0] = -inf, then top-k.  torch.topk leaves tie order unspecified; north_star fixes it to
    LOWEST INDEX FIRST, which is what a stable descending sort gives.  Returns (values, ids)."""
    s = np.array(scores, copy=True)
    if mask_col0:
        s[:, 0] = -np.inf
    order = np.argsort(-s, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(s, order, axis=1), order


def eval_metrics(topk_ids, pos_items, ks=(10, 20)):
    """Hit/NDCG/MRR@k with one positive per user ([upstream] RecBole metrics).
    Unrounded means (RecBole reports round(., 4))."""
    hit = (topk_ids == np.asarray(pos_items)[:, None])
    rank = np.arange(1, topk_ids.shape[1] + 1)
    out = {}
    for k in ks:
        h = hit[:, :k]
        out[f"hit@{k}"] = float(h.any(axis=1).mean())
        out[f"ndcg@{k}"] = float((h / np.log2(rank[:k] + 1)).sum(axis=1).mean())
        out[f"mrr@{k}"] = float((h / rank[:k]).sum(axis=1).mean())
    return out


def gated_recurrent_layer_bwd(x, p, grad_y, disable_conv1d=False):
    """Analytic backward of gated_recurrent_layer(literal_padding=True): gradients flow through the
    phantom steps exactly as autograd does in the reference.  Returns (dx, grads) with grads keyed by
    the module's parameter names."""
    B, T, D = x.shape
    C = p["Lambda"].shape[0]
    P = left_pad_len(T)
    Win, Wg, bg, Wo, Lam = p["input.weight"], p["gates.weight"], p["gates.bias"], p["output.weight"], p["Lambda"]
    cw, cb = p["conv1d.weight"].reshape(C, -1), p["conv1d.bias"]
    xz = x @ Win.T
    xs, z = xz[..., :C], xz[..., C:]
    xs_pad = np.concatenate([np.zeros((B, P, C), x.dtype), xs], axis=1)
    c = xs_pad if disable_conv1d else causal_conv1d_silu_fwd(xs_pad, cw, cb)
    g = c @ Wg.T + bg
    h = gated_scan_fwd(c, g[..., :C], g[..., C:], Lam)
    hT = h[:, P:]
    u = silu(z) * hT
    grads = {"output.weight": np.einsum("btd,btc->dc", grad_y, u)}
    du = grad_y @ Wo
    dz = du * hT * silu_grad(z)
    dh = np.concatenate([np.zeros((B, P, C), x.dtype), du * silu(z)], axis=1)
    dc, dr, di, dLam, _ = gated_scan_bwd(c, g[..., :C], g[..., C:], Lam, None, dh)
    dg = np.concatenate([dr, di], axis=-1)
    grads["Lambda"] = dLam
    grads["gates.weight"] = np.einsum("btg,btc->gc", dg, c)
    grads["gates.bias"] = dg.sum(axis=(0, 1))
    dc = dc + dg @ Wg
    if disable_conv1d:
        dxs_pad = dc
        grads["conv1d.weight"] = np.zeros_like(p["conv1d.weight"])
        grads["conv1d.bias"] = np.zeros_like(cb)
    else:
        dxs_pad, dcw, dcb = causal_conv1d_silu_bwd(xs_pad, cw, cb, dc)
        grads["conv1d.weight"] = dcw.reshape(p["conv1d.weight"].shape)
        grads["conv1d.bias"] = dcb
    dxz = np.concatenate([dxs_pad[:, P:], dz], axis=-1)
    grads["input.weight"] = np.einsum("btg,btd->gd", dxz, x)
    return dxz @ Win, grads
