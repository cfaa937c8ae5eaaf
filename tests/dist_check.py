"""Multi-GPU parity check, run as `python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1
tests/dist_check.py`: the row-sharded scoring and CE (NCCL exchange + CUDA kernels) must equal the single-table
kernels on the same data — ids identical, loss/grad within fp32 reduction-order noise."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from datamining_recblr_b200 import ops, sharded  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator().manual_seed(0)            # identical data on every rank
    B, N, D, k = 1000, 50021, 128, 10
    q = torch.randn(B, D, generator=g).to(torch.bfloat16).to(dev)
    table = (torch.randn(N, D, generator=g) * 0.3).to(torch.bfloat16).to(dev)
    pos = torch.randint(0, N, (B,), generator=g).to(dev)
    b = sharded.shard_bounds(N, world)
    shard = table[b[rank]:b[rank + 1]].contiguous()
    # scoring
    s1, i1 = ops.fullsort_topk(q, table, k, mask_id=0)
    s2, i2 = sharded.sharded_topk(q, shard, k, id_offset=b[rank], mask_id=0)
    assert torch.equal(i1, i2), "sharded top-k ids differ from the single-table result"
    assert torch.equal(s1, s2)
    # CE forward + backward
    qf = q.float().requires_grad_(True)
    tf = table.float().requires_grad_(True)
    l1 = ops.fullsort_cross_entropy(qf, tf, pos)
    l1.backward()
    qs = q.float().requires_grad_(True)
    ts = shard.float().requires_grad_(True)
    l2 = sharded.sharded_cross_entropy(qs, ts, pos, id_offset=b[rank])
    l2.backward()
    assert abs(float(l1) - float(l2)) <= 1e-6 * abs(float(l1)), (float(l1), float(l2))
    assert (qs.grad - qf.grad).abs().max() <= 2e-2 * qf.grad.abs().max()
    assert (ts.grad - tf.grad[b[rank]:b[rank + 1]]).abs().max() <= 2e-2 * tf.grad.abs().max()
    dist.barrier()
    if rank == 0:
        print(f"dist_check ok: world={world} loss={float(l1):.6f}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
