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


def _cpu_ce_grads(qb, eb, pos, lse, scale, id_offset=0):
    """Oracle stand-in for ops.fullsort_ce_grads: scale * (softmax - onehot) applied to both operands, GLOBAL lse."""
    p = torch.exp(qb.double() @ eb.double().T - lse.double()[:, None])
    loc = pos - id_offset
    own = (loc >= 0) & (loc < eb.shape[0])
    p[torch.arange(len(pos))[own], loc[own]] -= 1.0
    p *= scale
    return (p @ eb.double()).float(), (p.T @ qb.double()).float()


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
