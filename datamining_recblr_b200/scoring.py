"""Fused scoring heads for the reference's baseline models: SASRec and BERT4Rec end in the same
`Q . E^T` (+ `output_bias`) contraction followed by a full-softmax CE or RecBole's mask + top-k as RecBLR does
(sasrec.py:129-133,144-150; bert4rec.py:200-213,230-242), so they reuse the tcgen05 kernels of the RecBLR path.  These are
the two expressions a maintainer swaps inside those models' `calculate_loss` / evaluation; the models themselves
(transformer encoders from RecBole) are out of scope.  CUDA tensors only — there is no dense fallback here.
"""
import torch

from . import ops


def _rows(weight, n_items):
    # bert4rec.py:202,237: `item_embedding.weight[:n_items]` drops the [MASK] token row appended to the table
    return weight if n_items is None else weight[:n_items]


@torch.no_grad()
def full_sort_topk(seq_output, item_embedding_weight, k, n_items=None, output_bias=None, mask_padding_item=True):
    """(scores [B, k] fp32, ids [B, k] int64) of `seq_output @ weight[:n_items]^T (+ output_bias)` with RecBole's
    `scores[:, 0] = -inf` and lowest-id-first ties — SASRec: sasrec.py:144-150, BERT4Rec: bert4rec.py:230-242."""
    scores, ids = ops.fullsort_topk(seq_output, _rows(item_embedding_weight, n_items), k,
                                    mask_id=0 if mask_padding_item else -1, item_bias=output_bias)
    return scores, ids.long()


def cross_entropy(seq_output, item_embedding_weight, pos_items, n_items=None, output_bias=None, targets=None):
    """Full-softmax CE of the baseline models without materialising the logits.
    SASRec (sasrec.py:129-133): seq_output [B, H], pos_items [B] -> mean CE.
    BERT4Rec (bert4rec.py:200-213): seq_output [B, mask_len, H], pos_items [B, mask_len], targets = (masked_index > 0):
    `sum(CE * targets) / sum(targets)` = the mean CE over the positions with target 1, which are selected here before
    the kernels run (one boolean gather; the padded mask slots cost no GEMM work)."""
    H = seq_output.shape[-1]
    q, pos = seq_output.reshape(-1, H), pos_items.reshape(-1)
    if targets is not None:
        sel = targets.reshape(-1) > 0
        q, pos = q[sel], pos[sel]
    return ops.fullsort_cross_entropy(q, _rows(item_embedding_weight, n_items), pos, item_bias=output_bias)
