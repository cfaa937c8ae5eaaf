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
    # configs[4]: data-parallel model step with the CE row-sharded over the ranks == one process on the concatenated batch
    from datamining_recblr_b200.recblr import RecBLR
    from oracle.reference_loader import FakeDataset, make_config
    from oracle import torch_port as TP
    n_items, L, Bl = 3001, 50, 64
    torch.manual_seed(3)
    cfgs = {}
    for impl in ("fused", "sharded"):
        cfg = make_config(hidden_size=64, num_layers=2, dropout_prob=0.0, max_len=L, ce_impl=impl)
        cfg["device"] = dev
        cfgs[impl] = cfg
    torch.manual_seed(3)
    ref_model = RecBLR(cfgs["fused"], FakeDataset(n_items)).to(dev)
    dp_model = RecBLR(cfgs["sharded"], FakeDataset(n_items)).to(dev)
    dp_model.load_state_dict(ref_model.state_dict())
    seq, lens, tgt = TP.synthetic_batch(Bl * world, L, n_items, seed=9)
    full = {"item_id_list": seq.to(dev), "item_length": lens.to(dev), "item_id": tgt.to(dev)}
    mine = {k: v[rank * Bl:(rank + 1) * Bl] for k, v in full.items()}
    ref_loss = ref_model.calculate_loss(full)
    ref_loss.backward()
    dp_loss = dp_model.calculate_loss(mine)
    dp_loss.backward()
    sharded.allreduce_gradients(dp_model.parameters(), average=False)
    assert abs(float(ref_loss) - float(dp_loss)) <= 1e-5 * abs(float(ref_loss)), (float(ref_loss), float(dp_loss))
    for (n1, p1), (n2, p2) in zip(ref_model.named_parameters(), dp_model.named_parameters()):
        den = p1.grad.abs().max().clamp_min(1e-12)
        assert (p1.grad - p2.grad).abs().max() <= 3e-2 * den, (n1, float((p1.grad - p2.grad).abs().max()), float(den))
    # H7: the tied table itself row-sharded (fp32 master shard + replicated bf16 copy): TWO optimizer steps on `world`
    # ranks == two steps of one process on the concatenated batches (table rounded to bf16 so both read the same values)
    import copy
    torch.manual_seed(5)
    cfg = make_config(hidden_size=64, num_layers=2, dropout_prob=0.0, max_len=L)
    cfg["device"] = dev
    one = RecBLR(cfg, FakeDataset(n_items)).to(dev)
    with torch.no_grad():
        one.item_embedding.weight.copy_(one.item_embedding.weight.to(torch.bfloat16).float())
    shd = copy.deepcopy(one)
    sit = sharded.shard_item_table(shd)
    opt1 = torch.optim.SGD(one.parameters(), lr=0.5)
    opt2 = torch.optim.SGD(shd.parameters(), lr=0.5)
    sit.attach(opt2)
    for step in range(2):
        seq, lens, tgt = TP.synthetic_batch(Bl * world, L, n_items, seed=20 + step)
        full = {"item_id_list": seq.to(dev), "item_length": lens.to(dev), "item_id": tgt.to(dev)}
        mine = {k: v[rank * Bl:(rank + 1) * Bl] for k, v in full.items()}
        opt1.zero_grad(set_to_none=True)
        opt2.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            l_one = one.calculate_loss(full)
            l_shd = shd.calculate_loss(mine)
        l_one.backward()
        l_shd.backward()
        sharded.allreduce_gradients(sharded.dense_parameters(shd), average=False)
        assert abs(float(l_one) - float(l_shd)) <= 2e-3 * abs(float(l_one)), (step, float(l_one), float(l_shd))
        g1 = one.item_embedding.weight.grad[sit.lo:sit.hi]
        g2 = sit.master.grad[:sit.n_local]
        den_t = one.item_embedding.weight.grad.abs().max()
        assert (g1 - g2).abs().max() <= 3e-2 * den_t, ("table grad", step, float((g1 - g2).abs().max()), float(den_t))
        d1 = {n: p.grad for n, p in one.named_parameters() if n != "item_embedding.weight"}
        d2 = {n: p.grad for n, p in shd.named_parameters() if n != "item_embedding.weight"}
        for n in d1:
            den = d1[n].abs().max().clamp_min(1e-12)
            assert (d1[n] - d2[n]).abs().max() <= 5e-2 * den, (n, step, float((d1[n] - d2[n]).abs().max()), float(den))
        opt1.step()
        opt2.step()            # post-hook: refresh of the replicated bf16 copy (cast + in-place all-gather)
        with torch.no_grad():  # keep the single-process table bf16-representable like the sharded copy
            full_w = sit.full_weight()
            assert (full_w - one.item_embedding.weight).abs().max() <= 0.5 * 3e-2 * float(den_t) + 1e-6
            one.item_embedding.weight.copy_(full_w.to(torch.bfloat16).float())
            sit.master[:sit.n_local].copy_(one.item_embedding.weight[sit.lo:sit.hi])
            sit.refresh()
    # data-parallel eval through the sharded table (gather users -> per-shard kernel -> one all-gather -> in-place merge ->
    # this rank's slice) == the single-table kernel on the replicated copy for the SAME model's seq_output: identical ids
    shd.eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        s_shd, i_shd = shd.full_sort_topk(mine, 10)
        q_mine = shd.forward(mine["item_id_list"], mine["item_length"])
    s_one, i_one = ops.fullsort_topk(q_mine, sit.table_bf16[:sit.n_items], 10, mask_id=0)
    assert torch.equal(i_shd, i_one.long()) and torch.equal(s_shd, s_one), "sharded eval differs from the single table"
    dist.barrier()
    if rank == 0:
        print(f"dist_check ok: world={world} loss={float(l1):.6f} dp_sharded_ce_loss={float(dp_loss):.6f} "
              f"sharded_table_loss={float(l_shd):.6f}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
