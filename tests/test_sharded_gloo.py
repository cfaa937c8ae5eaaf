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


def _oracle_local_topk(q, shard, k, mask_id=0, id_offset=0):
    s = q.double().numpy() @ shard.double().numpy().T
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


def test_shard_bounds_cover_table():
    from datamining_recblr_b200 import sharded
    for n, g in ((10, 3), (1000003, 8), (5, 8)):
        b = sharded.shard_bounds(n, g)
        assert b[0] == 0 and b[-1] == n and all(x <= y for x, y in zip(b, b[1:]))
