"""Multi-GPU forms of the full-sort scoring and full-softmax CE: the item table is ROW-SHARDED over the ranks of a
`torch.distributed` process group (rank g owns global item ids [bounds[g], bounds[g+1]); the pad row 0 lives on rank 0).
New functionality — the reference is single-GPU — whose oracle is the single-table result.

Only small tensors cross NVLink: queries are replicated by the caller (data parallel eval gathers them first),
per-shard top-k lists [B, k] are all-gathered and merged by (score desc, id asc); CE statistics are combined with one
MAX and two SUM all-reduces of [B] vectors; dE stays shard-local, dQ partials are all-reduced.

The per-shard arithmetic runs in the CUDA kernels (`ops`); the exchange logic below is backend-agnostic so that the
world_size-2 `gloo` tests can drive it on CPU with the oracle standing in for the kernels (`local_*` hooks).
"""
import torch
import torch.distributed as dist

from . import ops


def shard_bounds(n_items, world):
    """Row ranges of the item table: rank g owns [b[g], b[g+1])."""
    return [n_items * g // world for g in range(world + 1)]


def _world(group):
    return dist.get_world_size(group) if dist.is_initialized() else 1


def sharded_topk(q, table_shard, k, id_offset, mask_id=0, group=None, local_topk=None, local_merge=None, item_bias=None):
    """Top-k of q @ table^T (+ item_bias) over the whole (sharded) table.  q [B, D] must be identical on every rank;
    `item_bias` is this rank's slice [rows of table_shard] of the bias vector (bert4rec.py:230-242).
    Returns (scores [B, k] fp32, ids [B, k] int32), identical on every rank and to the single-table result."""
    local_topk = local_topk or ops.fullsort_topk
    local_merge = local_merge or ops.topk_merge
    kw = {} if item_bias is None else {"item_bias": item_bias}
    s, i = local_topk(q, table_shard, k, mask_id=mask_id, id_offset=id_offset, **kw)
    world = _world(group)
    if world == 1:
        return s, i
    B = q.shape[0]
    cs = torch.empty((world * B, k), dtype=torch.float32, device=s.device)   # rank-major concatenation
    ci = torch.empty((world * B, k), dtype=torch.int32, device=s.device)
    dist.all_gather_into_tensor(cs, s.contiguous(), group=group)
    dist.all_gather_into_tensor(ci, i.contiguous(), group=group)
    cs = cs.view(world, B, k).permute(1, 0, 2).reshape(B, world * k).contiguous()
    ci = ci.view(world, B, k).permute(1, 0, 2).reshape(B, world * k).contiguous()
    return local_merge(cs, ci, k)


def combine_ce_stats(row_max, row_sum, pos_logit, group=None):
    """Global logsumexp and positive logit from per-shard (max, sum exp(. - max), pos_logit-or-0):
        m = max_g m_g;  s = sum_g s_g * exp(m_g - m);  lse = m + log s;  pos = sum_g pos_g."""
    if _world(group) == 1:
        return row_max + torch.log(row_sum), pos_logit
    m = row_max.clone()
    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    s = row_sum * torch.exp(row_max - m)
    packed = torch.stack([s, pos_logit])
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return m + torch.log(packed[0]), packed[1]


class _ShardedCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, table_shard, pos, id_offset, group, reduce_dq, item_bias):
        qb, eb = ops._bf16_rows(q), ops._bf16_rows(table_shard)
        if item_bias is not None:   # the bias rides the GEMM as 64 extra K columns (ops._augment_with_bias)
            qb, eb = ops._augment_with_bias(qb, eb, item_bias)
        m, s, pl = ops.fullsort_ce_stats(qb, eb, pos, id_offset=id_offset)
        lse, pl = combine_ce_stats(m, s, pl, group)
        ctx.save_for_backward(qb, eb, pos, lse)
        ctx.meta = (id_offset, group, q.dtype, table_shard.dtype, reduce_dq, q.shape[1],
                    None if item_bias is None else item_bias.dtype)
        return (lse - pl).mean()

    @staticmethod
    def backward(ctx, grad_loss):
        qb, eb, pos, lse = ctx.saved_tensors
        id_offset, group, qd, ed, reduce_dq, D, bd = ctx.meta
        dQ, dE = ops.fullsort_ce_grads(qb, eb, pos, lse, 1.0 / qb.shape[0], id_offset=id_offset)
        dQ = dQ[:, :D].contiguous()          # no-op without a bias; drops the gradient of the constant 1-columns
        if reduce_dq and _world(group) > 1:  # every shard contributes P_shard E_shard to dQ; dE is shard-local
            dist.all_reduce(dQ, op=dist.ReduceOp.SUM, group=group)
        g = grad_loss.float()
        dbias = None if bd is None else (dE[:, D] * g).to(bd)
        return (dQ * g).to(qd), (dE[:, :D] * g).to(ed), None, None, None, None, dbias


def sharded_cross_entropy(q, table_shard, pos, id_offset, group=None, reduce_dq=True, item_bias=None):
    """Mean full-softmax CE over the whole sharded table; q [B, D] and pos [B] (GLOBAL ids) identical on every rank.
    Gradients: dtable_shard is this rank's rows; dq is the full gradient on every rank (one all-reduce) unless
    reduce_dq=False, in which case it is this shard's partial and the caller sums it (e.g. the reduce-scatter in the
    backward of an autograd-aware all-gather).  `item_bias` = this rank's slice of a per-item logit bias
    (bert4rec.py:200-213); its gradient is shard-local like the table's."""
    return _ShardedCE.apply(q, table_shard, ops._check_pos(pos, q.shape[0]), int(id_offset), group, bool(reduce_dq),
                            item_bias)


# ----------------------------------------------------------------------------- data parallel + sharded CE (configs[4])
def data_parallel_sharded_ce(seq_output, table, pos_items, group=None):
    """Loss of a DATA-PARALLEL step whose full-softmax CE is SHARDED by item rows (BASELINE.json configs[4]).

    Every rank holds the whole (replicated) item table `table [n_items, D]` — it needs it for the input gather — but
    scores only its row shard: the per-rank seq_output [B, D] / pos_items [B] are all-gathered (autograd-aware), every
    rank computes the CE statistics and gradients of ALL G*B users against rows [lo, hi) with the fused kernels, and
    the result is the GLOBAL mean loss (identical on every rank).  Backward: the all-gather's backward sums the per-shard
    dQ partials over ranks and hands each rank its own users' slice; dE lands in rows [lo, hi) of `table.grad` only.
    Because the loss is already the global mean, gradients must be SUMMED over ranks afterwards
    (`allreduce_gradients(params, average=False)`): for the table that sum also assembles the per-shard CE rows."""
    world = _world(group)
    if world == 1:
        return ops.fullsort_cross_entropy(seq_output, table, pos_items)
    from torch.distributed.nn.functional import all_gather
    rank = dist.get_rank(group)
    q_all = torch.cat(all_gather(seq_output.contiguous(), group=group), dim=0)
    pos_list = [torch.empty_like(pos_items) for _ in range(world)]
    dist.all_gather(pos_list, pos_items.contiguous(), group=group)
    b = shard_bounds(table.shape[0], world)
    # the all_gather's backward reduce-scatters the per-shard dQ partials: no extra all-reduce inside the CE
    return sharded_cross_entropy(q_all, table[b[rank]:b[rank + 1]], torch.cat(pos_list), id_offset=b[rank], group=group,
                                 reduce_dq=False)


def allreduce_gradients(params, group=None, average=True):
    """One flat NCCL all-reduce over the gradients of `params` (mean for an ordinary data-parallel loss, sum when the
    loss is already a global mean as in data_parallel_sharded_ce)."""
    if _world(group) == 1:
        return
    ps = [p for p in params if p.grad is not None]
    flat = torch.cat([p.grad.reshape(-1) for p in ps])
    dist.all_reduce(flat, op=dist.ReduceOp.AVG if average else dist.ReduceOp.SUM, group=group)
    off = 0
    for p in ps:  # the reduced gradients are handed back as VIEWS of the flat buffer: no copy-back kernels
        p.grad = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
