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


def sharded_topk(q, table_shard, k, id_offset, mask_id=0, group=None, local_topk=None, local_merge=None, item_bias=None,
                 gathered_merge=None):
    """Top-k of q @ table^T (+ item_bias) over the whole (sharded) table.  q [B, D] must be identical on every rank;
    `item_bias` is this rank's slice [rows of table_shard] of the bias vector (bert4rec.py:230-242).
    Returns (scores [B, k] fp32, ids [B, k] int32), identical on every rank and to the single-table result.

    Exchange: the local kernel writes its (scores | ids) lists into the two halves of ONE packed [2, B, k] buffer, a
    single all-gather moves them, and the merge kernel reads the rank-major result in place (no second collective, no
    permute / contiguous copies); every call in the sequence is capturable in a CUDA graph."""
    local_topk = local_topk or ops.fullsort_topk
    kw = {} if item_bias is None else {"item_bias": item_bias}
    world = _world(group)
    if world == 1:
        return local_topk(q, table_shard, k, mask_id=mask_id, id_offset=id_offset, **kw)
    B = q.shape[0]
    packed = torch.empty((2, B, k), dtype=torch.float32, device=q.device)
    out = (packed[0], packed[1].view(torch.int32))
    if local_topk is ops.fullsort_topk:
        local_topk(q, table_shard, k, mask_id=mask_id, id_offset=id_offset, out=out, **kw)
    else:   # test stand-ins return fresh tensors
        s, i = local_topk(q, table_shard, k, mask_id=mask_id, id_offset=id_offset, **kw)
        out[0].copy_(s)
        out[1].copy_(i)
    gathered = torch.empty((world * 2, B, k), dtype=torch.float32, device=q.device)   # concatenation along dim 0
    dist.all_gather_into_tensor(gathered, packed, group=group)
    gathered = gathered.view(world, 2, B, k)
    if local_merge is None and gathered_merge is None:
        return ops.topk_merge_gathered(gathered, k)
    if gathered_merge is not None:
        return gathered_merge(gathered, k)
    cs = gathered[:, 0].permute(1, 0, 2).reshape(B, world * k).contiguous()           # CPU test path
    ci = gathered[:, 1].view(torch.int32).permute(1, 0, 2).reshape(B, world * k).contiguous()
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
        dQ, dE = ops.fullsort_ce_grads(qb, eb, pos, lse, 1.0 / qb.shape[0], id_offset=id_offset, scale_dev=grad_loss)
        dQ = dQ[:, :D].contiguous()          # no-op without a bias; drops the gradient of the constant 1-columns
        if reduce_dq and _world(group) > 1:  # every shard contributes P_shard E_shard to dQ; dE is shard-local
            dist.all_reduce(dQ, op=dist.ReduceOp.SUM, group=group)
        dbias = None if bd is None else dE[:, D].to(bd)
        return dQ.to(qd), dE[:, :D].to(ed), None, None, None, None, dbias


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


# ----------------------------------------------------------------------------- row-sharded TIED item table (training)
def _reduce_scatter_rows(full, group):
    """Sum of `full` [world * n, ...] over ranks, this rank keeping its n rows (reduce-scatter; gloo, which has no
    reduce-scatter, all-reduces and slices — CPU tests only)."""
    world, rank = _world(group), dist.get_rank(group)
    n = full.shape[0] // world
    if dist.get_backend(group) == "nccl":
        out = torch.empty((n, *full.shape[1:]), dtype=full.dtype, device=full.device)
        dist.reduce_scatter_tensor(out, full.contiguous(), op=dist.ReduceOp.SUM, group=group)
        return out
    full = full.clone()
    dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)
    return full[rank * n:(rank + 1) * n].clone()


class _Kernels:
    """The per-shard CUDA entry points the sharded table calls; the gloo CPU tests substitute oracle stand-ins."""
    embed_fwd = None   # filled lazily (ops needs the library)


class _TiedEmbedLN(torch.autograd.Function):
    """LayerNorm(dropout(table[ids])) reading the REPLICATED bf16 copy; its backward does not return a table gradient
    but routes the per-token row gradients to the row owners' fp32 gradient shards (ShardedItemTable.add_row_grads)."""

    @staticmethod
    def forward(ctx, ids, gamma, beta, sit, eps, p, seed, seed_dev):
        out, saved = sit.k.embed_fwd(ids, sit.table_bf16, gamma, beta, eps, p, seed, seed_dev)
        ctx.sit, ctx.saved, ctx.args = sit, saved, (ids, gamma, eps, p, seed, seed_dev)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        sit = ctx.sit
        ids, gamma, eps, p, seed, seed_dev = ctx.args
        dgamma, dbeta = sit.k.embed_bwd(sit, ids, gamma, grad_out, ctx.saved, p, seed, seed_dev)
        return None, dgamma.to(gamma.dtype), dbeta.to(gamma.dtype), None, None, None, None, None


class _TiedCE(torch.autograd.Function):
    """Global-mean full-softmax CE of a data-parallel batch against the row-sharded table.  Forward: all-gather the
    queries; per-shard reference maximum -> all-reduce MAX; ONE fused exponential pass per shard giving its share of the
    softmax denominator and of the unnormalised dQ (ops.fullsort_ce_fwd_dq); tiny all-reduce SUM.  Backward: the dQ shares
    are normalised and reduce-scattered back to the users' ranks, and the dE pass writes STRAIGHT into the owner's
    gradient shard."""

    @staticmethod
    def forward(ctx, q, pos, sit):
        world, group = sit.world, sit.group
        qb = q.detach().to(torch.bfloat16).contiguous()
        pos = pos.contiguous()
        if world > 1:
            q_all = torch.empty((world * qb.shape[0], qb.shape[1]), dtype=qb.dtype, device=qb.device)
            dist.all_gather_into_tensor(q_all, qb, group=group)
            pos_all = torch.empty(world * pos.shape[0], dtype=pos.dtype, device=pos.device)
            dist.all_gather_into_tensor(pos_all, pos, group=group)
        else:
            q_all, pos_all = qb, pos
        shard = sit.shard_bf16()
        ref = sit.k.ce_rowmax(q_all, shard)
        if world > 1:
            dist.all_reduce(ref, op=dist.ReduceOp.MAX, group=group)
        acc, s = sit.k.ce_fwd_dq(q_all, shard, ref)
        pl, prow = ops.pos_logits(q_all, shard, pos_all, sit.lo)
        if world > 1:
            packed = torch.stack([s, pl])
            dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
            s, pl = packed[0], packed[1]
        sit.k.assert_finite(s)
        lse = ref + torch.log(s)
        ctx.sit, ctx.q_dtype = sit, q.dtype
        ctx.save_for_backward(q_all, pos_all, lse, acc, s, prow)
        return (lse - pl).mean()

    @staticmethod
    def backward(ctx, grad_loss):
        sit = ctx.sit
        q_all, pos_all, lse, acc, s, prow = ctx.saved_tensors
        scale = grad_loss.float() / q_all.shape[0]
        dq_part = (acc / s[:, None] - prow) * scale   # this shard's share of dL/dq for ALL users of the global batch
        de = sit.grad_buffer()                        # fp32 [rows_per, D]; rows [0, n_local) are OVERWRITTEN by the kernel
        sit.k.ce_de(q_all, sit.shard_bf16(), pos_all, lse, 1.0 / q_all.shape[0], sit.lo, grad_loss, de[:sit.n_local])
        sit.master.grad = de
        sit._ce_written = True
        dq = _reduce_scatter_rows(dq_part, sit.group) if sit.world > 1 else dq_part
        return dq.to(ctx.q_dtype), None, None


class ShardedItemTable:
    """The tied item-embedding table of RecBLR (RecBLR.py:39, used by the input gather :76 AND the CE :99-103) ROW-SHARDED
    over the ranks of a process group for training at configs[4] scale:

      master      fp32 [rows_per, D] nn.Parameter — this rank's rows [lo, hi) (+ zero pad rows up to rows_per, so that every
                  shard has the same size); the optimizer and its state exist for these rows only
      table_bf16  bf16 [world * rows_per, D], replicated — what the input gather and the tensor-core CE read; refreshed
                  after every optimizer step from the masters (one cast of the shard + one in-place all-gather)

    Per step nothing of size [n_items, D] is all-reduced: the CE's dE is complete on the owner (every rank scores ALL
    users of the global batch against its rows) and is written by the kernel directly into `master.grad`; the input
    gather's row gradients travel as (ids, bf16 rows) of the local tokens, all-gathered and scatter-added by each owner
    for its id range; the loss is the GLOBAL mean, so dense parameters are afterwards SUMMED over ranks
    (`allreduce_gradients(dense, average=False)`).  world == 1 degenerates to the same code without collectives."""

    def __init__(self, full_weight, group=None, padding_idx=0, kernels=None):
        self.group = group
        self.world = _world(group)
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        n_items, D = full_weight.shape
        self.n_items, self.D, self.padding_idx = n_items, D, padding_idx
        self.rows_per = -(-n_items // self.world)
        self.lo = min(self.rank * self.rows_per, n_items)
        self.hi = min(self.lo + self.rows_per, n_items)
        self.n_local = self.hi - self.lo
        dev = full_weight.device
        master = torch.zeros((self.rows_per, D), dtype=torch.float32, device=dev)
        master[:self.n_local] = full_weight.detach()[self.lo:self.hi].float()
        self.master = torch.nn.Parameter(master)
        self.table_bf16 = torch.zeros((self.world * self.rows_per, D), dtype=torch.bfloat16, device=dev)
        self._grad = None
        self._ce_written = False
        assert self.n_local >= 1, f"item table of {n_items} rows cannot be sharded {self.world} ways"
        self.k = kernels or _cuda_kernels()
        self.refresh()

    # ---- views
    def shard_bf16(self):
        return self.table_bf16[self.lo:self.lo + self.n_local]   # lo == rank * rows_per: global id == row of the copy

    def grad_buffer(self):
        if self._grad is None:
            self._grad = torch.zeros_like(self.master)   # pad rows stay zero for ever
        return self._grad

    # ---- the two uses of the table
    def embed_layernorm(self, ids, gamma, beta, eps, dropout_p, seed, seed_dev):
        return _TiedEmbedLN.apply(ids, gamma, beta, self, float(eps), float(dropout_p), int(seed), seed_dev)

    def cross_entropy(self, q, pos):
        return _TiedCE.apply(q, pos, self)

    def add_row_grads(self, ids, drows):
        """Owner-side accumulation of the input gather's gradient: ids int64 [n], drows [n, D] of THIS rank's tokens."""
        assert self._ce_written, "sharded table: the CE backward must run before the embedding backward (same loss)"
        de = self.master.grad
        if self.world > 1:
            ids_all = torch.empty(self.world * ids.numel(), dtype=ids.dtype, device=ids.device)
            dist.all_gather_into_tensor(ids_all, ids.reshape(-1).contiguous(), group=self.group)
            rows_all = torch.empty((self.world * drows.shape[0], drows.shape[1]), dtype=drows.dtype, device=drows.device)
            dist.all_gather_into_tensor(rows_all, drows.contiguous(), group=self.group)
        else:
            ids_all, rows_all = ids.reshape(-1), drows
        self.k.scatter_rows(ids_all, rows_all, de, self.lo, self.hi, self.padding_idx)
        self._ce_written = False

    # ---- after the optimizer step
    @torch.no_grad()
    def refresh(self):
        """bf16 copy <- masters: cast this rank's rows into its slice, then one in-place all-gather of the slices."""
        mine = self.table_bf16[self.rank * self.rows_per:(self.rank + 1) * self.rows_per]
        mine.copy_(self.master)
        if self.world > 1:
            dist.all_gather_into_tensor(self.table_bf16, mine, group=self.group)

    def attach(self, optimizer):
        """Refresh the replicated copy automatically after every `optimizer.step()`."""
        optimizer.register_step_post_hook(lambda *_: self.refresh())
        return self

    @torch.no_grad()
    def full_weight(self):
        """fp32 [n_items, D] assembled from the masters (checkpointing; one all-gather)."""
        if self.world == 1:
            return self.master.detach()[:self.n_items].clone()
        out = torch.empty((self.world * self.rows_per, self.D), dtype=torch.float32, device=self.master.device)
        dist.all_gather_into_tensor(out, self.master.detach().contiguous(), group=self.group)
        return out[:self.n_items].clone()

    @torch.no_grad()
    def full_sort_topk(self, q, k, mask_id=0):
        """Data-parallel eval: every rank holds different users; gather them, score all against the local rows, merge
        the per-shard lists, keep this rank's users."""
        qb = q.detach().to(torch.bfloat16).contiguous()
        if self.world > 1:
            q_all = torch.empty((self.world * qb.shape[0], qb.shape[1]), dtype=qb.dtype, device=qb.device)
            dist.all_gather_into_tensor(q_all, qb, group=self.group)
        else:
            q_all = qb
        s, i = sharded_topk(q_all, self.shard_bf16(), k, id_offset=self.lo, mask_id=mask_id, group=self.group)
        n = qb.shape[0]
        return s[self.rank * n:(self.rank + 1) * n], i[self.rank * n:(self.rank + 1) * n]


def _cuda_kernels():
    """The CUDA entry points behind ShardedItemTable (bf16 table, bf16 activations)."""
    from . import _lib as L

    class K:
        @staticmethod
        def embed_fwd(ids, table, gamma, beta, eps, p, seed, seed_dev):
            n_items, D = table.shape
            ids_c = ids.contiguous()
            gf, bf = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
            n = ids_c.numel()
            out = torch.empty((*ids.shape, D), dtype=table.dtype, device=table.device)
            mean = torch.empty(n, dtype=torch.float32, device=table.device)
            rstd = torch.empty_like(mean)
            L.check(L.load().bdlru_embed_ln_fwd(L.ptr(ids_c), L.ptr(table), L.ptr(gf), L.ptr(bf), L.ptr(out), L.ptr(mean),
                                                L.ptr(rstd), n, n_items, D, eps, p, seed, L.ptr(seed_dev),
                                                L.dtype_tag(table), L.dtype_tag(out), L.stream_ptr(table)))
            return out, (ids_c, gf, mean, rstd)

        @staticmethod
        def embed_bwd(sit, ids, gamma, grad_out, saved, p, seed, seed_dev):
            ids_c, gf, mean, rstd = saved
            table = sit.table_bf16
            n_items, D = table.shape
            n = ids_c.numel()
            grad_out = grad_out.to(table.dtype).contiguous()
            dgamma = torch.empty(D, dtype=torch.float32, device=table.device)
            dbeta = torch.empty_like(dgamma)
            lib = L.load()
            nws = lib.bdlru_embed_ln_bwd_workspace_bytes(n, D)
            ws = ops._workspace(table.device, nws)
            tail = (n, n_items, D, p, seed, L.ptr(seed_dev), sit.padding_idx, L.dtype_tag(table), L.dtype_tag(grad_out),
                    L.stream_ptr(table))
            if sit.world == 1:   # the owner is this rank: scatter straight into the gradient shard the CE already wrote
                assert sit._ce_written, "sharded table: the CE backward must run before the embedding backward"
                L.check(lib.bdlru_embed_ln_bwd(L.ptr(ids_c), L.ptr(table), L.ptr(gf), L.ptr(grad_out), L.ptr(mean),
                                               L.ptr(rstd), L.ptr(sit.master.grad), L.ptr(dgamma), L.ptr(dbeta), L.ptr(ws),
                                               nws, *tail))
                sit._ce_written = False
            else:
                drows = torch.empty((n, D), dtype=table.dtype, device=table.device)
                L.check(lib.bdlru_embed_ln_bwd_rows(L.ptr(ids_c), L.ptr(table), L.ptr(gf), L.ptr(grad_out), L.ptr(mean),
                                                    L.ptr(rstd), L.ptr(drows), L.ptr(dgamma), L.ptr(dbeta), L.ptr(ws), nws,
                                                    *tail))
                sit.add_row_grads(ids_c, drows)
            return dgamma, dbeta

        @staticmethod
        def ce_rowmax(q_all, shard):
            return ops.fullsort_rowmax(q_all, shard, ops.CE_REFERENCE_STRIDE)

        @staticmethod
        def ce_fwd_dq(q_all, shard, ref):
            return ops.fullsort_ce_fwd_dq(q_all, shard, ref)

        @staticmethod
        def assert_finite(s):
            torch._assert_async(torch.isfinite(s).all(), "sharded CE: the sampled softmax reference is > 80 below a row "
                                "maximum (exp overflow): set ops.CE_REFERENCE_STRIDE = 1")

        @staticmethod
        def ce_de(q_all, shard, pos_all, lse, scale, lo, grad_loss, out_de):
            ops.fullsort_ce_grads(q_all, shard, pos_all, lse, scale, id_offset=lo, scale_dev=grad_loss, out_de=out_de,
                                  want_dq=False)

        @staticmethod
        def scatter_rows(ids_all, rows_all, dst, lo, hi, padding_idx):
            ops.scatter_add_rows(ids_all, rows_all, dst, lo, hi, padding_idx)

    return K


class ShardedTableOptimizer:
    """Adam for a model whose item table is row-sharded: torch.optim.Adam (fused) for the dense parameters, and for the
    table ONE kernel (`bdlru_table_adam_step`) that applies the same Adam update to this rank's fp32 master rows AND writes
    their bf16 copy, followed by the in-place all-gather of the copy.  Drop-in for the optimizer object RecBole's trainer
    holds (`zero_grad`, `step`, `state_dict`, `param_groups`): construct it instead of `torch.optim.Adam(model.parameters())`.
    Same arithmetic as torch.optim.Adam(amsgrad=False) — tests compare the two step by step."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, kernel=None):
        self.sit = model.table_shard
        assert self.sit is not None, "call sharded.shard_item_table(model) first"
        dense = dense_parameters(model)
        self.dense = torch.optim.Adam(dense, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                                      fused=dense[0].is_cuda, capturable=dense[0].is_cuda) if dense else None
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        m = self.sit.master
        self.exp_avg = torch.zeros_like(m)
        self.exp_avg_sq = torch.zeros_like(m)
        self.steps = 0
        self._kernel = kernel or _table_adam_cuda

    @property
    def param_groups(self):
        return (self.dense.param_groups if self.dense is not None else []) + [
            {"params": [self.sit.master], "lr": self.lr, "betas": self.betas, "eps": self.eps}]

    def zero_grad(self, set_to_none=True):
        if self.dense is not None:
            self.dense.zero_grad(set_to_none=set_to_none)
        self.sit.master.grad = None if set_to_none else self.sit.master.grad

    @torch.no_grad()
    def step(self):
        if self.dense is not None:
            self.dense.step()
        sit = self.sit
        if sit.master.grad is not None:
            self.steps += 1
            mine = sit.table_bf16[sit.rank * sit.rows_per:(sit.rank + 1) * sit.rows_per]
            self._kernel(sit.master.data, sit.master.grad, self.exp_avg, self.exp_avg_sq, mine, self.lr, self.betas[0],
                         self.betas[1], self.eps, self.weight_decay, self.steps)
            if sit.world > 1:
                dist.all_gather_into_tensor(sit.table_bf16, mine, group=sit.group)

    def state_dict(self):
        return {"dense": self.dense.state_dict() if self.dense is not None else None, "steps": self.steps,
                "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq}

    def load_state_dict(self, sd):
        if self.dense is not None and sd["dense"] is not None:
            self.dense.load_state_dict(sd["dense"])
        self.steps = int(sd["steps"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])


def _table_adam_cuda(p, g, m, v, p_bf16, lr, b1, b2, eps, wd, step):
    from . import _lib as L
    L.require_cuda(p, g, m, v, p_bf16)
    assert p.is_contiguous() and g.is_contiguous() and m.is_contiguous() and v.is_contiguous() and p_bf16.is_contiguous()
    assert g.dtype == torch.float32 and p_bf16.dtype == torch.bfloat16 and p_bf16.numel() == p.numel()
    L.check(L.load().bdlru_table_adam_step(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), L.ptr(p_bf16), p.numel(), float(lr),
                                           float(b1), float(b2), float(eps), float(wd), int(step), L.stream_ptr(p)))


def shard_item_table(model, group=None):
    """Switches a constructed RecBLR (identical full table on every rank: same seed, or a loaded checkpoint) to the
    row-sharded tied table: `model.item_embedding.weight` becomes this rank's fp32 master shard (what the optimizer
    sees), `model.table_shard` the ShardedItemTable that `_front` / `calculate_loss` / `full_sort_topk` go through.
    Call before building the optimizer, then `model.table_shard.attach(optimizer)`."""
    full = model.item_embedding.weight
    sit = ShardedItemTable(full.data, group=group, padding_idx=0)
    del model.item_embedding.weight
    model.item_embedding.weight = sit.master        # registered as the module's parameter again (nn.Parameter)
    model.table_shard = sit
    return sit


def dense_parameters(model):
    """Parameters replicated on every rank (everything but the sharded table): the ones to all-reduce."""
    sit = getattr(model, "table_shard", None)
    return [p for p in model.parameters() if sit is None or p is not sit.master]
