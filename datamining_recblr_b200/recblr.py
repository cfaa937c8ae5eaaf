"""Host-side mirror of the reference model file (RecBLR.py) over the sm_100a kernels.

Same class names, constructor arguments, config keys, parameter names/shapes (checkpoints of the reference
load unchanged: SURVEY.md §8b) and the same four entry points RecBole's trainer calls
(`forward`, `calculate_loss`, `predict`, `full_sort_predict`), so `run.py` switches to this path by
importing `RecBLR` from here instead of from the reference's `RecBLR.py`.

What differs is only HOW `GatedRecurrentLayer.forward` and the front end are evaluated:
  * no left pad to a power of two (RecBLR.py:177-179, 203-204): the P phantom steps the reference feeds
    through the conv bias are folded into the batch-independent initial state h0[C] (closed form,
    differentiable), so results are identical while 28 % fewer steps are computed;
  * conv + SiLU, gate math + scan + z-gate, and gather + dropout + LayerNorm are one kernel each,
    all on the native channel-last [B, T, C] layout (no transposes);
  * `full_sort_topk` / the CE branch of `calculate_loss` use the fused tcgen05 kernels that never write the
    [B, n_items] logits (`full_sort_predict` still returns the dense matrix RecBole's stock trainer needs).

There is no CPU path: the ops raise on non-CUDA tensors.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from . import ops

try:  # RecBole is the host framework of the reference (requirements.txt:3); absent in the build image
    from recbole.model.abstract_recommender import SequentialRecommender
    from recbole.model.loss import BPRLoss
except ImportError:  # the attributes RecBLR.py:20,37-38,83,87-92 rely on
    class SequentialRecommender(nn.Module):
        def __init__(self, config, dataset):
            super().__init__()
            self.USER_ID = config["USER_ID_FIELD"]
            self.ITEM_ID = config["ITEM_ID_FIELD"]
            self.ITEM_SEQ = self.ITEM_ID + config["LIST_SUFFIX"]
            self.ITEM_SEQ_LEN = config["ITEM_LIST_LENGTH_FIELD"]
            self.POS_ITEM_ID = self.ITEM_ID
            self.NEG_ITEM_ID = config["NEG_PREFIX"] + self.ITEM_ID
            self.max_seq_length = config["MAX_ITEM_LIST_LENGTH"]
            self.n_items = dataset.num(self.ITEM_ID)
            self.device = config["device"]

        def gather_indexes(self, output, gather_index):
            index = gather_index.view(-1, 1, 1).expand(-1, -1, output.shape[-1])
            return output.gather(dim=1, index=index).squeeze(1)

    class BPRLoss(nn.Module):
        def __init__(self, gamma=1e-10):
            super().__init__()
            self.gamma = gamma

        def forward(self, pos_score, neg_score):
            return -torch.log(self.gamma + torch.sigmoid(pos_score - neg_score)).mean()


def softplus_inverse(x):
    return torch.log(torch.exp(x) - 1)  # the reference's exact expression (RecBLR.py:14-15): bit-identical fresh inits


def _autocast_bf16():
    """torch.bfloat16 under a bf16 autocast region (activations then leave the front end in bf16), else None."""
    if torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16:
        return torch.bfloat16
    return None


def _cfg(config, key, default=None):
    """RecBole's Config returns None for unknown keys; plain dicts raise — accept both."""
    try:
        v = config[key]
    except KeyError:
        v = None
    return default if v is None else v


class RecBLR(SequentialRecommender):
    """RecBLR.py:18-122.  Extra (optional) config keys, all defaulting to the fast path:
        ce_impl      'fused' (tcgen05 online-softmax CE, bf16 operands / fp32 accumulate) | 'dense' (fp32 logits)
                     | 'sharded' (data parallel step, CE row-sharded over torch.distributed's default group: the loss is
                       the GLOBAL mean, sum gradients over ranks afterwards — sharded.data_parallel_sharded_ce)
        fused_front  True: gather+dropout+LayerNorm in one kernel
    """

    def __init__(self, config, dataset):
        super().__init__(config, dataset)
        self.hidden_size = config["hidden_size"]
        self.loss_type = config["loss_type"]
        self.num_layers = config["num_layers"]
        self.dropout_prob = config["dropout_prob"]
        self.expand = config["expand"]
        self.d_conv = config["d_conv"]
        self.bd_lru_only = _cfg(config, "bd_lru_only")
        self.disable_conv1d = _cfg(config, "disable_conv1d")
        self.disable_ffn = _cfg(config, "disable_ffn")
        if self.bd_lru_only:  # RecBLR.py:33-35
            self.disable_conv1d = True
            self.disable_ffn = True
        self.ce_impl = _cfg(config, "ce_impl", "fused")
        self.fused_front = bool(_cfg(config, "fused_front", True))

        self.item_embedding = nn.Embedding(self.n_items, self.hidden_size, padding_idx=0)
        self.layer_norm = nn.LayerNorm(self.hidden_size, eps=1e-12)
        self.dropout = nn.Dropout(self.dropout_prob)
        self.recurrent_layers = nn.ModuleList([
            RecurrentLayer(d_model=self.hidden_size, d_conv=self.d_conv, expand=self.expand,
                           dropout=self.dropout_prob, num_layers=self.num_layers, bd_lru_only=self.bd_lru_only,
                           disable_conv1d=self.disable_conv1d, disable_ffn=self.disable_ffn)
            for _ in range(self.num_layers)
        ])
        if self.loss_type == "BPR":
            self.loss_fct = BPRLoss()
        elif self.loss_type == "CE":
            self.loss_fct = nn.CrossEntropyLoss()
        else:
            raise NotImplementedError("Make sure 'loss_type' in ['BPR', 'CE']!")
        self.apply(self._init_weights)
        # step counter of the fused front end's dropout stream, ON THE DEVICE so that a training step captured in a
        # CUDA graph draws a new mask at every replay (not a parameter, not saved in checkpoints)
        self.register_buffer("_dropout_step", torch.zeros(1, dtype=torch.int64), persistent=False)
        # set by sharded.shard_item_table(model): the tied table row-sharded over a process group
        self.table_shard = None

    def _init_weights(self, module):  # RecBLR.py:66-73 (the pad row 0 is re-randomised too: SURVEY quirk 2)
        if isinstance(module, (nn.Linear, nn.Embedding)):
            module.weight.data.normal_(std=0.02)
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)
        if isinstance(module, nn.Linear) and module.bias is not None:
            module.bias.data.zero_()

    # ------------------------------------------------------------------ RecBLR.py:75-84
    def _front(self, item_seq):
        p = self.dropout_prob if self.training else 0.0
        D = self.hidden_size
        if self.table_shard is not None:   # gather from the replicated bf16 copy; row gradients go to the row owners
            out = self.table_shard.embed_layernorm(item_seq, self.layer_norm.weight, self.layer_norm.bias,
                                                   self.layer_norm.eps, p, self._seed_base(),
                                                   self._dropout_step if p > 0.0 else None)
            return out if _autocast_bf16() is not None else out.float()
        if self.fused_front and D % 4 == 0 and D <= 512:
            seed_dev = self._dropout_step if p > 0.0 else None   # advanced once per training forward (RecBLR.forward)
            seed = self._seed_base()
            return ops.embed_layernorm(item_seq, self.item_embedding.weight, self.layer_norm.weight,
                                       self.layer_norm.bias, eps=self.layer_norm.eps, dropout_p=p, seed=seed,
                                       padding_idx=0, seed_dev=seed_dev,
                                       out_dtype=_autocast_bf16())
        return self.layer_norm(self.dropout(self.item_embedding(item_seq)))

    def forward(self, item_seq, item_seq_len):
        dropping = self.training and self.dropout_prob > 0
        if dropping:
            # one tick per training forward, whichever front-end path runs: every fused dropout site of this step
            # (front end, residual LayerNorms, FFN) derives its mask from (seed, this counter, site)
            self._dropout_step.add_(1)
        item_emb = self._front(item_seq)
        ctx = (self._seed_base(), self._dropout_step if dropping else None)
        last = len(self.recurrent_layers) - 1
        for i, layer in enumerate(self.recurrent_layers):
            if i == last and not self.training and not torch.is_grad_enabled():
                # only position len-1 of the last layer's output is ever used (RecBLR.py:83): the recurrence still walks the
                # whole sequence, but the out-projection, both residual LayerNorms and the FFN (RecBLR.py:140-145) run
                # on [B, D] instead of [B, L, D] and the gather is folded in
                return layer.forward_last(item_emb, item_seq_len - 1)
            item_emb = layer(item_emb, dropout_ctx=(ctx[0] + 7919 * (i + 1), ctx[1]))
        return self.gather_indexes(item_emb, item_seq_len - 1)

    @staticmethod
    def _seed_base():
        """Host part of the fused dropouts' seed: torch's global seed, decorrelated across data-parallel ranks (every rank
        usually calls manual_seed with the same value, but must not drop the same (row, channel) positions)."""
        rank = 0
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            rank = torch.distributed.get_rank()
        return ((torch.initial_seed() + 0x632BE59BD9B4E019 * rank) * 0x9E3779B97F4A7C15) & 0x3FFFFFFFFFFFFFFF

    # ------------------------------------------------------------------ RecBLR.py:86-103
    def calculate_loss(self, interaction):
        item_seq = interaction[self.ITEM_SEQ]
        item_seq_len = interaction[self.ITEM_SEQ_LEN]
        seq_output = self.forward(item_seq, item_seq_len)
        pos_items = interaction[self.POS_ITEM_ID]
        if self.table_shard is not None:
            if self.loss_type != "CE":
                raise NotImplementedError("the row-sharded item table supports loss_type 'CE' only")
            return self.table_shard.cross_entropy(seq_output, pos_items)   # GLOBAL-mean loss: sum dense grads over ranks
        if self.loss_type == "BPR":
            neg_items = interaction[self.NEG_ITEM_ID]
            pos_score = torch.sum(seq_output * self.item_embedding(pos_items), dim=-1)
            neg_score = torch.sum(seq_output * self.item_embedding(neg_items), dim=-1)
            return self.loss_fct(pos_score, neg_score)
        table = self.item_embedding.weight
        if self.ce_impl == "sharded" and ops.fullsort_supported(self.hidden_size):
            from .sharded import data_parallel_sharded_ce
            return data_parallel_sharded_ce(seq_output, table, pos_items)
        if self.ce_impl in ("fused", "sharded") and ops.fullsort_supported(self.hidden_size):
            return ops.fullsort_cross_entropy(seq_output, table, pos_items)
        logits = torch.matmul(seq_output, table.transpose(0, 1))
        return self.loss_fct(logits, pos_items)

    # ------------------------------------------------------------------ RecBLR.py:105-122
    def predict(self, interaction):
        if self.table_shard is not None:
            raise NotImplementedError("predict() needs the whole table: use full_sort_topk with the sharded table")
        seq_output = self.forward(interaction[self.ITEM_SEQ], interaction[self.ITEM_SEQ_LEN])
        test_item_emb = self.item_embedding(interaction[self.ITEM_ID])
        return torch.mul(seq_output, test_item_emb).sum(dim=1)

    def full_sort_predict(self, interaction):
        if self.table_shard is not None:
            raise NotImplementedError("the dense [B, n_items] scores do not exist with a sharded table: use full_sort_topk")
        seq_output = self.forward(interaction[self.ITEM_SEQ], interaction[self.ITEM_SEQ_LEN])
        return torch.matmul(seq_output, self.item_embedding.weight.transpose(0, 1))

    @torch.no_grad()
    def full_sort_topk(self, interaction, k, mask_padding_item=True):
        """Fused form of `full_sort_predict` + RecBole's `scores[:, 0] = -inf; torch.topk(scores, k)`
        (SURVEY §3.5): returns (scores [B, k] fp32, ids [B, k] int64) ordered by score descending with the
        LOWEST item id first among ties, without materialising [B, n_items]."""
        seq_output = self.forward(interaction[self.ITEM_SEQ], interaction[self.ITEM_SEQ_LEN])
        if self.table_shard is not None:
            scores, ids = self.table_shard.full_sort_topk(seq_output, k, mask_id=0 if mask_padding_item else -1)
            return scores, ids.long()
        scores, ids = ops.fullsort_topk(seq_output, self.item_embedding.weight, k,
                                        mask_id=0 if mask_padding_item else -1)
        return scores, ids.long()


def _linear(layer, x):
    """nn.Linear with bias evaluated through ops.linear_bias (same math; fast bias gradient) on CUDA tensors."""
    if layer.bias is not None:
        return ops.linear_bias(x, layer.weight, layer.bias)   # raises on CPU tensors: there is no CPU path
    return layer(x)


def _dropout_seed(owner, dropout_ctx, site):
    """(host seed, device step counter or None) of one fused-dropout call site."""
    if dropout_ctx is not None and dropout_ctx[1] is not None:
        return dropout_ctx[0] + 104729 * site, dropout_ctx[1]
    owner._host_step = getattr(owner, "_host_step", 0) + 1  # standalone layer: advance a host-side step count
    return (RecBLR._seed_base() + 104729 * site + owner._host_step) & 0x3FFFFFFFFFFFFFFF, None


def _silu_dropout(owner, dropout, x, dropout_ctx, site):
    """dropout(silu(x)) (RecBLR.py:219-221) through the fused kernel."""
    p = dropout.p if owner.training else 0.0
    if x.numel() % 8 != 0:  # shape the kernel does not take: ATen on the same device
        return dropout(F.silu(x))
    seed, seed_dev = _dropout_seed(owner, dropout_ctx, site) if p > 0.0 else (0, None)
    return ops.silu_dropout(x, dropout_p=p, seed=seed, seed_dev=seed_dev)


def _residual_ln(owner, norm, dropout, x, residual, dropout_ctx, site):
    """LayerNorm(dropout(x) + residual) (RecBLR.py:142 / 221-225) through the fused kernel."""
    D = x.shape[-1]
    p = dropout.p if owner.training else 0.0
    if D % 4 != 0 or D > 512:  # shape the kernel does not take: ATen on the same device
        return norm(dropout(x) + residual)
    seed, seed_dev = _dropout_seed(owner, dropout_ctx, site) if p > 0.0 else (0, None)
    return ops.add_dropout_layernorm(x, residual, norm.weight, norm.bias, eps=norm.eps, dropout_p=p, seed=seed,
                                     seed_dev=seed_dev)


class RecurrentLayer(nn.Module):
    """RecBLR.py:124-145."""

    def __init__(self, d_model, d_conv, expand, dropout, num_layers, bd_lru_only, disable_conv1d, disable_ffn):
        super().__init__()
        self.num_layers = num_layers
        self.disable_ffn = disable_ffn
        self.behavior_modeling = GatedRecurrentLayer(d_model=d_model, expansion_factor=expand, kernel_size=d_conv,
                                                     bd_lru_only=bd_lru_only, disable_conv1d=disable_conv1d)
        self.dropout = nn.Dropout(dropout)
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-12)
        self.ffn = FeedForward(d_model=d_model, inner_size=d_model * 4, dropout=dropout)

    # Inference: the whole first half as ONE tcgen05 kernel (csrc/fused_layer.cu) when the shape is the one it is built for
    # and the problem is small enough to be launch-bound (measured on B200: at 4 096 x 50 tokens the eager eval forward
    # drops from 0.95 to 0.80 ms; in steady-state throughput its three serialised MMA phases per 32-step tile make it 0.86-0.90x
    # of "cuBLAS projections + fused core kernel + add_ln", so larger problems keep that path).
    fused_layer = True
    fused_layer_max_tokens = 1 << 18

    def forward(self, input_tensor, dropout_ctx=None):
        """dropout_ctx = (seed, device step counter or None): the counter stream shared by the model's fused dropouts
        (RecBLR.forward supplies it; standalone use falls back to a host-side step count)."""
        bm = self.behavior_modeling
        if (self.fused_layer and not self.training and not torch.is_grad_enabled()
                and input_tensor.shape[0] * input_tensor.shape[1] <= self.fused_layer_max_tokens
                and ops.bdlru_layer_supported(input_tensor, bm.gates.weight.shape[1], bm.conv1d.weight.shape[-1])):
            # in-projection, conv, gates GEMM, recurrence, z-gate, out-projection, residual and LayerNorm: ONE kernel
            hidden_states = ops.bdlru_layer_fused(
                input_tensor, bm.input.weight, bm.conv1d.weight.squeeze(1), bm.conv1d.bias, bm.gates.weight, bm.gates.bias,
                bm.Lambda, bm.phantom_state(input_tensor.shape[1]), bm.output.weight, self.layer_norm.weight,
                self.layer_norm.bias, self.layer_norm.eps, use_conv=not bm.disable_conv1d)
            if not self.disable_ffn:
                hidden_states = self.ffn(hidden_states, dropout_ctx)
            return hidden_states
        # the layer input feeds the in-projection AND the residual: taken from one autograd node (ops.linear_tap)
        hidden_states, residual = self.behavior_modeling(input_tensor, return_tap=True)
        hidden_states = _residual_ln(self, self.layer_norm, self.dropout, hidden_states, residual, dropout_ctx, 1)
        if not self.disable_ffn:
            hidden_states = self.ffn(hidden_states, dropout_ctx)
        return hidden_states

    @torch.no_grad()
    def forward_last(self, input_tensor, last_index):
        """Inference form of `forward` followed by `gather_indexes(., last_index)`: [B, L, D], [B] -> [B, D].  Same
        arithmetic per kept row; nothing downstream of the scan is computed for the L-1 rows that would be discarded."""
        bm = self.behavior_modeling
        _, seq_len, _ = input_tensor.shape
        xz = bm.input(input_tensor)
        y = bm.core(xz, seq_len)
        y_last = _gather_rows(y, last_index)                       # [B, C]
        x_last = _gather_rows(input_tensor, last_index)            # [B, D] residual
        h = self.layer_norm(bm.output(y_last) + x_last.to(y_last.dtype))
        if not self.disable_ffn:
            f = self.ffn
            h = f.layer_norm(f.w_2(F.silu(f.w_1(h))) + h)
        return h


def _gather_rows(x, index):
    """x[b, index[b], :] for x [B, T, C] -> [B, C] (SequentialRecommender.gather_indexes)."""
    return x.gather(1, index.view(-1, 1, 1).expand(-1, 1, x.shape[-1])).squeeze(1)


class GatedRecurrentLayer(nn.Module):
    """RecBLR.py:148-207 — the BD-LRU layer."""

    def __init__(self, d_model=64, expansion_factor=2, kernel_size=4, bd_lru_only=False, disable_conv1d=False):
        super().__init__()
        self.bd_lru_only = bd_lru_only
        self.disable_conv1d = disable_conv1d
        r_min, r_max = 0.9, 0.999  # exp(-softplus(Lambda)) spans [0.9, 0.999] at init (RecBLR.py:153-158)
        lo = softplus_inverse(torch.tensor(-math.log(r_min))).item()
        hi = softplus_inverse(torch.tensor(-math.log(r_max))).item()
        hidden = int(d_model * expansion_factor)
        self.input = nn.Linear(d_model, 2 * hidden, bias=False)
        self.conv1d = nn.Conv1d(in_channels=hidden, out_channels=hidden, bias=True, kernel_size=kernel_size,
                                groups=hidden, padding=kernel_size - 1)
        self.gates = nn.Linear(hidden, 2 * hidden, bias=True)
        self.Lambda = nn.Parameter(torch.linspace(lo, hi, hidden))
        self.output = nn.Linear(hidden, d_model, bias=False)
        self.fused_core = True   # inference: one tcgen05 kernel for conv + gates GEMM + recurrence (ops.bdlru_core_fused)

    def phantom_state(self, seq_len):
        """State the reference's left zero-pad leaves in front of the first real step (RecBLR.py:177-199):
        on the P = 2^ceil(log2 T) - T padded steps the conv emits silu(conv bias) =: s, so
            h0 = b' (1 - a^P) / (1 - a),   a, b' = gate math at the constant input s.
        Batch independent, [C], fp32, differentiable w.r.t. conv1d.bias, gates.*, Lambda.  None when P == 0
        or the conv is disabled (then padded steps are exactly zero)."""
        pad_len = 2 ** ((seq_len - 1).bit_length()) - seq_len
        if pad_len == 0 or self.disable_conv1d:
            return None
        # one fused kernel each way (as torch ops this is ~45 tiny kernels per layer and step); CUDA only
        return ops.phantom_h0(self.conv1d.bias, self.gates.weight, self.gates.bias, self.Lambda, pad_len)

    def core(self, xz, seq_len):
        """conv -> gates -> recurrence -> z-gate on xz = (x | z): the fused tcgen05 inference kernel when nothing needs a
        gradient and the shape is the one it is built for (bf16, C = 128, conv width 4), else the separate kernels."""
        h0 = self.phantom_state(seq_len)
        conv_w = self.conv1d.weight.squeeze(1)
        if (not torch.is_grad_enabled() and self.fused_core and conv_w.shape[1] == 4
                and ops.bdlru_core_supported(xz, xz.shape[-1] // 2)):
            return ops.bdlru_core_fused(xz, conv_w, self.conv1d.bias, self.gates.weight, self.gates.bias, self.Lambda,
                                        h0=h0, use_conv=not self.disable_conv1d)
        return ops.bdlru_block(xz, conv_w, self.conv1d.bias, self.gates.weight, self.gates.bias, self.Lambda, h0=h0,
                               use_conv=not self.disable_conv1d)

    def forward(self, x, return_tap=False):
        """return_tap=True also returns x as a second output of the in-projection's autograd node (for the caller's
        residual connection)."""
        _, seq_len, _ = x.shape
        # [B, T, 2C] = (x | z); consumed in place by the kernels, no chunk copies
        if return_tap:
            xz, x = ops.linear_tap(x, self.input.weight)
        else:
            xz = self.input(x)
        y = self.core(xz, seq_len)
        return (self.output(y), x) if return_tap else self.output(y)


class FeedForward(nn.Module):
    """RecBLR.py:210-227."""

    def __init__(self, d_model, inner_size, dropout=0.2):
        super().__init__()
        self.w_1 = nn.Linear(d_model, inner_size)
        self.w_2 = nn.Linear(inner_size, d_model)
        self.dropout = nn.Dropout(dropout)
        self.layer_norm = nn.LayerNorm(d_model, eps=1e-12)

    def forward(self, input_tensor, dropout_ctx=None):
        hidden_states, residual = ops.linear_tap(input_tensor, self.w_1.weight, self.w_1.bias)
        hidden_states = _silu_dropout(self, self.dropout, hidden_states, dropout_ctx, 3)
        hidden_states = _linear(self.w_2, hidden_states)
        return _residual_ln(self, self.layer_norm, self.dropout, hidden_states, residual, dropout_ctx, 2)
