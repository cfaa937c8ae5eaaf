"""world_size-2 `gloo` tests (CPU) of the multi-GPU exchange logic in datamining_recblr_b200/sharded.py: row-sharded
top-k with all-gather + merge, and the (max, sumexp) combination of the sharded CE.  The per-shard arithmetic is done
by the numpy oracle through the module's `local_*` hooks (the CUDA kernels need a GPU; their sharded form is checked
on the device by tests/test_gpu_fullsort.py and tests/dist_check.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bdlru_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_local_topk(q, shard, k, mask_id=0, id_offset=0, item_bias=None):
    s = q.double().numpy() @ shard.double().numpy().T
    if item_bias is not None:
        s = s + item_bias.double().numpy()[None, :]
    ids = np.arange(shard.shape[0]) + id_offset
    if mask_id >= 0:
        s[:, ids == mask_id] = -np.inf
    order = np.argsort(-s, axis=1, kind="stable")[:, :k]
    vals = np.take_along_axis(s, order, axis=1)
    gid = ids[order]
    gid = np.where(np.isneginf(vals), -1, gid)
    return torch.tensor(vals, dtype=torch.float32), torch.tensor(gid, dtype=torch.int32)


def _oracle_merge(cs, ci, k):
    s, i = cs.double().numpy(), ci.numpy().astype(np.int64)
    key_i = np.where(i < 0, np.iinfo(np.int64).max, i)
    order = np.lexsort((key_i, -s), axis=1)[:, :k]
    return (torch.tensor(np.take_along_axis(s, order, 1), dtype=torch.float32),
            torch.tensor(np.take_along_axis(i, order, 1), dtype=torch.int32))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from datamining_recblr_b200 import sharded
    rng = np.random.default_rng(0)      # same data on every rank
    B, N, D, k = 37, 1001, 16, 10
    q = torch.tensor(rng.integers(-2, 3, size=(B, D)), dtype=torch.float32)      # exact ties on purpose
    table = torch.tensor(rng.integers(-1, 2, size=(N, D)), dtype=torch.float32)
    pos = torch.tensor(rng.integers(0, N, size=B))
    b = sharded.shard_bounds(N, world)
    shard = table[b[rank]:b[rank + 1]]
    s, i = sharded.sharded_topk(q, shard, k, id_offset=b[rank], mask_id=0, local_topk=_oracle_local_topk,
                                local_merge=_oracle_merge)
    # CE statistics of this shard (oracle), combined across ranks by the product code
    logits = (q.double() @ shard.double().T).numpy()
    m = logits.max(1)
    ssum = np.exp(logits - m[:, None]).sum(1)
    pl = np.zeros(B)
    own = (pos.numpy() >= b[rank]) & (pos.numpy() < b[rank + 1])
    pl[own] = logits[np.arange(B)[own], pos.numpy()[own] - b[rank]]
    lse, plg = sharded.combine_ce_stats(torch.tensor(m), torch.tensor(ssum), torch.tensor(pl))
    if rank == 0:
        torch.save(dict(s=s, i=i, lse=lse, pl=plg, q=q, table=table, pos=pos), out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_topk_and_ce_combine_equal_single_table(tmp_path, world):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    r = torch.load(out)
    scores = r["q"].double().numpy() @ r["table"].double().numpy().T
    v_ref, i_ref = O.topk_lowest_index(scores, 10)
    assert (r["i"].numpy() == i_ref).all()
    assert np.allclose(r["s"].numpy(), v_ref)
    loss_ref, lse_ref, _, _ = O.ce_loss(r["q"].double().numpy(), r["table"].double().numpy(), r["pos"].numpy())
    assert np.abs(r["lse"].numpy() - lse_ref).max() < 1e-9
    assert abs(float((r["lse"] - r["pl"]).mean()) - loss_ref) < 1e-9


# ----------------------------------------------------------------------------- sharded CE autograd (with item bias)
def _cpu_ce_stats(qb, eb, pos, id_offset=0, item_bias=None):
    """Oracle stand-in for ops.fullsort_ce_stats on CPU (same contract: per-shard max, sum exp(. - max), pos logit or 0)."""
    assert item_bias is None                  # the product code augments the operands before this call
    logits = qb.double() @ eb.double().T
    m = logits.max(1).values
    ssum = torch.exp(logits - m[:, None]).sum(1)
    loc = pos - id_offset
    own = (loc >= 0) & (loc < eb.shape[0])
    pl = torch.zeros_like(m)
    pl[own] = logits[torch.arange(len(pos))[own], loc[own]]
    return m.float(), ssum.float(), pl.float()


def _cpu_ce_grads(qb, eb, pos, lse, scale, id_offset=0, scale_dev=None, out_dq=None, out_de=None, want_dq=True,
                  want_de=True):
    """Oracle stand-in for ops.fullsort_ce_grads: scale * (softmax - onehot) applied to both operands, GLOBAL lse."""
    p = torch.exp(qb.double() @ eb.double().T - lse.double()[:, None])
    loc = pos - id_offset
    own = (loc >= 0) & (loc < eb.shape[0])
    p[torch.arange(len(pos))[own], loc[own]] -= 1.0
    p *= scale * (1.0 if scale_dev is None else float(scale_dev))
    dq, de = (p @ eb.double()).float(), (p.T @ qb.double()).float()
    if out_de is not None:
        out_de.copy_(de)
        de = out_de
    return dq, de


def _ce_worker(rank, world, port, outdir, with_bias):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from datamining_recblr_b200 import ops, sharded
    ops.fullsort_ce_stats, ops.fullsort_ce_grads = _cpu_ce_stats, _cpu_ce_grads   # kernels -> oracle (test only)
    rng = np.random.default_rng(1)      # same data on every rank; values exact in bf16
    B, N, D = 23, 301, 64
    q = torch.tensor(rng.integers(-2, 3, size=(B, D)) * 0.25, dtype=torch.float32, requires_grad=True)
    table = torch.tensor(rng.integers(-1, 2, size=(N, D)) * 0.5, dtype=torch.float32)
    bias = torch.tensor(rng.integers(-8, 9, size=N) * 0.125, dtype=torch.float32)
    pos = torch.tensor(rng.integers(0, N, size=B))
    b = sharded.shard_bounds(N, world)
    shard = table[b[rank]:b[rank + 1]].clone().requires_grad_(True)
    bshard = bias[b[rank]:b[rank + 1]].clone().requires_grad_(True) if with_bias else None
    loss = sharded.sharded_cross_entropy(q, shard, pos, id_offset=b[rank], item_bias=bshard)
    (loss * 2.0).backward()
    torch.save(dict(loss=loss.detach(), dq=q.grad, dshard=shard.grad, dbias=None if bshard is None else bshard.grad,
                    q=q.detach(), table=table, bias=bias, pos=pos, lo=b[rank], hi=b[rank + 1]),
               os.path.join(outdir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("with_bias", [False, True])
def test_sharded_cross_entropy_autograd_equals_single_table(tmp_path, with_bias):
    """The autograd wrapper of the row-sharded CE (statistics combine, dQ all-reduce, shard-local dE / d(bias), operand
    augmentation for the bias) on 2 gloo ranks, with the oracle standing in for the two CUDA entry points."""
    world = 2
    mp.spawn(_ce_worker, args=(world, _free_port(), str(tmp_path), with_bias), nprocs=world, join=True)
    rs = [torch.load(str(tmp_path / f"r{r}.pt")) for r in range(world)]
    q, table, pos = (rs[0][k].double().numpy() for k in ("q", "table", "pos"))
    pos = pos.astype(np.int64)
    bias = rs[0]["bias"].double().numpy() if with_bias else None
    loss_ref, _, dQ, dE = O.ce_loss(q, table, pos, bias)
    for r in rs:
        assert abs(float(r["loss"]) - loss_ref) < 1e-5 * abs(loss_ref)
        assert np.abs(r["dq"].numpy() / 2 - dQ).max() < 1e-5 * np.abs(dQ).max()          # full gradient on every rank
        assert np.abs(r["dshard"].numpy() / 2 - dE[r["lo"]:r["hi"]]).max() < 1e-5 * np.abs(dE).max()
        assert r["dq"].shape == (23, 64) and r["dshard"].shape == (r["hi"] - r["lo"], 64)
        if with_bias:
            db = O.ce_bias_grad(q, table, pos, bias)
            assert np.abs(r["dbias"].numpy() / 2 - db[r["lo"]:r["hi"]]).max() < 1e-5 * np.abs(db).max()


def _bias_topk_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from datamining_recblr_b200 import sharded
    rng = np.random.default_rng(2)
    B, N, D, k = 19, 503, 16, 10
    q = torch.tensor(rng.integers(-2, 3, size=(B, D)), dtype=torch.float32)
    table = torch.tensor(rng.integers(-1, 2, size=(N, D)), dtype=torch.float32)
    bias = torch.tensor(rng.integers(-3, 4, size=N), dtype=torch.float32)      # integer scores: exact ties across shards
    b = sharded.shard_bounds(N, world)
    s, i = sharded.sharded_topk(q, table[b[rank]:b[rank + 1]], k, id_offset=b[rank], mask_id=0,
                                local_topk=_oracle_local_topk, local_merge=_oracle_merge,
                                item_bias=bias[b[rank]:b[rank + 1]])
    if rank == 0:
        torch.save(dict(s=s, i=i, q=q, table=table, bias=bias), out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_topk_with_item_bias_equals_single_table(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_bias_topk_worker, args=(3, _free_port(), out), nprocs=3, join=True)
    r = torch.load(out)
    scores = O.full_sort_scores(r["q"].double().numpy(), r["table"].double().numpy(), r["bias"].double().numpy())
    v_ref, i_ref = O.topk_lowest_index(scores, 10)
    assert (r["i"].numpy() == i_ref).all() and np.allclose(r["s"].numpy(), v_ref)


def test_shard_bounds_cover_table():
    from datamining_recblr_b200 import sharded
    for n, g in ((10, 3), (1000003, 8), (5, 8)):
        b = sharded.shard_bounds(n, g)
        assert b[0] == 0 and b[-1] == n and all(x <= y for x, y in zip(b, b[1:]))


# ----------------------------------------------------------------------------- row-sharded TIED table (training, H7)
class _CpuTableKernels:
    """Oracle stand-ins (torch float64 on CPU) for the five CUDA entry points behind sharded.ShardedItemTable."""

    @staticmethod
    def embed_fwd(ids, table, gamma, beta, eps, p, seed, seed_dev):
        assert p == 0.0
        with torch.enable_grad():     # Function.forward runs under no_grad; the stand-in differentiates with autograd
            rows = table[ids.reshape(-1)].double().requires_grad_(True)
            g, b = gamma.detach().double().requires_grad_(True), beta.detach().double().requires_grad_(True)
            out = torch.nn.functional.layer_norm(rows, (table.shape[1],), g, b, eps)
        return out.detach().float().view(*ids.shape, -1), (ids, rows, g, b, out)

    @staticmethod
    def embed_bwd(sit, ids, gamma, grad_out, saved, p, seed, seed_dev):
        ids, rows, g, b, out = saved
        drows, dg, db = torch.autograd.grad(out, (rows, g, b), grad_out.double().reshape(out.shape))
        drows[ids.reshape(-1) == sit.padding_idx] = 0
        sit.add_row_grads(ids.reshape(-1), drows.float())
        return dg.float(), db.float()

    @staticmethod
    def ce_rowmax(q_all, shard):
        # a deliberately SLOPPY reference (true maximum minus 3): the result must not depend on it (shift invariance)
        return ((q_all.double() @ shard.double().T).max(1).values - 3.0).float()

    @staticmethod
    def ce_fwd_dq(q_all, shard, ref):
        p = torch.exp(q_all.double() @ shard.double().T - ref.double()[:, None])
        return (p @ shard.double()).float(), p.sum(1).float()

    @staticmethod
    def assert_finite(s):
        assert torch.isfinite(s).all()

    @staticmethod
    def ce_de(q_all, shard, pos_all, lse, scale, lo, grad_loss, out_de):
        _cpu_ce_grads(q_all, shard, pos_all, lse, scale, id_offset=lo, scale_dev=grad_loss, out_de=out_de)

    @staticmethod
    def scatter_rows(ids_all, rows_all, dst, lo, hi, padding_idx):
        own = (ids_all >= lo) & (ids_all < hi) & (ids_all != padding_idx)
        dst.index_add_(0, ids_all[own] - lo, rows_all[own].float())


def _tied_data(world, N=53, D=16, B=6, Lq=5):
    rng = np.random.default_rng(4)
    table = torch.tensor(rng.integers(-4, 5, size=(N, D)) * 0.125, dtype=torch.float32)     # exact in bf16
    gamma = torch.tensor(1.0 + rng.integers(-2, 3, size=D) * 0.125, dtype=torch.float32)
    beta = torch.tensor(rng.integers(-2, 3, size=D) * 0.125, dtype=torch.float32)
    ids = [torch.tensor(rng.integers(0, N, size=(B, Lq))) for _ in range(world)]           # includes padding id 0
    pos = [torch.tensor(rng.integers(1, N, size=B)) for _ in range(world)]
    return table, gamma, beta, ids, pos


def _tied_worker(rank, world, port, outdir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from datamining_recblr_b200 import sharded
    table, gamma, beta, ids, pos = _tied_data(world)
    sit = sharded.ShardedItemTable(table, padding_idx=0, kernels=_CpuTableKernels)
    g, b = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    x = sit.embed_layernorm(ids[rank], g, b, 1e-12, 0.0, 0, None)          # [B, Lq, D]
    q = torch.tanh(x.double()).mean(1).float()                             # stand-in for the recurrent layers
    q = q + (q.to(torch.bfloat16).float() - q).detach()                    # the CE rounds q to bf16 (a real cast would also
                                                                           # round the GRADIENT to bf16 on the way back)
    loss = sit.cross_entropy(q, pos[rank])
    (loss * 3.0).backward()
    dgb = torch.stack([g.grad, b.grad])
    if world > 1:
        dist.all_reduce(dgb)                                               # dense parameters: summed over ranks
    grad = sit.master.grad.clone()
    with torch.no_grad():
        sit.master -= 0.25 * sit.master.grad                               # a "step", then refresh the replicated copy
    sit.refresh()
    torch.save(dict(loss=loss.detach(), grad=grad, dgb=dgb, lo=sit.lo, hi=sit.hi, n_local=sit.n_local,
                    copy=sit.table_bf16.clone(), full=sit.full_weight()), os.path.join(outdir, f"r{rank}.pt"))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2, 3])
def test_sharded_tied_table_step_equals_single_table_autograd(tmp_path, world):
    """ShardedItemTable on `world` gloo ranks (different users per rank), the oracle standing in for the kernels: the
    GLOBAL-mean loss, the per-owner gradient shards (CE part written in place + input-gather rows exchanged and
    scatter-added, padding id skipped), the summed dense gradients and the refreshed replicated bf16 copy all equal plain
    float64 autograd over ONE tied table with the concatenated batch."""
    if world == 1:
        _tied_worker(0, 1, 0, str(tmp_path))
    else:
        mp.spawn(_tied_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rs = [torch.load(str(tmp_path / f"r{r}.pt")) for r in range(world)]
    table, gamma, beta, ids, pos = _tied_data(world)
    T = table.double().requires_grad_(True)
    g, b = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    all_ids, all_pos = torch.cat(ids), torch.cat(pos)
    emb = torch.nn.functional.embedding(all_ids, T, padding_idx=0)
    x = torch.nn.functional.layer_norm(emb, (T.shape[1],), g, b, 1e-12)
    q = torch.tanh(x.float().double()).mean(1)
    q = q + (q.float().to(torch.bfloat16).double() - q).detach()           # same bf16 rounding, identity gradient
    loss = torch.nn.functional.cross_entropy(q @ T.T, all_pos)
    (loss * 3.0).backward()
    for r in rs:
        assert abs(float(r["loss"]) - float(loss)) < 1e-5 * abs(float(loss))
        want = T.grad[r["lo"]:r["hi"]]
        got = r["grad"][:r["n_local"]].double()
        assert (got - want).abs().max() < 2e-5 * T.grad.abs().max(), (r["lo"], (got - want).abs().max())
        assert float(r["grad"][r["n_local"]:].abs().sum()) == 0.0         # pad rows of the last shard stay zero
        assert (r["dgb"][0].double() - g.grad).abs().max() < 1e-4 * g.grad.abs().max()
        assert (r["dgb"][1].double() - b.grad).abs().max() < 1e-4 * b.grad.abs().max()
        new_full = (table.double() - 0.25 * T.grad).float()
        assert (r["full"] - new_full).abs().max() < 1e-5
        assert torch.equal(r["copy"][:table.shape[0]], r["full"].to(torch.bfloat16))
        assert torch.equal(r["copy"], rs[0]["copy"])                       # replicated copy identical on every rank
