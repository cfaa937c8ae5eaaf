"""bench.py --workload score1m|score10m: synthetic full-sort scoring (BASELINE.json configs[3]).

4 096 users x N items, D = 128, bf16 operands / fp32 accumulate, top-10 with RecBole's column-0 mask.  The item table is
row-sharded over the ranks (rank g owns rows [g*N/G, (g+1)*N/G)); every rank scores all users against its shard with
the fused tcgen05 GEMM + streaming top-k kernel, the [B, k] candidate lists are exchanged with one NCCL all-gather and
merged per user by (score desc, id asc) — identical to the 1-GPU result by construction.  A step scores
the 4 096 users once; value = users / step time (fixed catalog => strong scaling in N).
"""
import json
import os
import statistics
import time

import torch


def run_score(args):
    import torch.distributed as dist
    from bench import SCORE_WORKLOADS, ClockSampler, cpu_score_baseline, dist_env, finish_distributed, ncu_traffic, peaks
    from datamining_recblr_b200 import _lib, ops, sharded
    from datamining_recblr_b200.timing import flush_l2

    rank, world, local = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N, D, B, k = SCORE_WORKLOADS[args.workload], 128, 4096, 10
    lo, hi = N * rank // world, N * (rank + 1) // world
    g = torch.Generator(device=dev).manual_seed(2020 + rank)
    # E ~ N(0, 0.02^2) like _init_weights (RecBLR.py:68); generated per shard on the device (no dataset, no network)
    E = (torch.randn(hi - lo, D, generator=g, device=dev) * 0.02).to(torch.bfloat16)
    gq = torch.Generator().manual_seed(2020)
    n_q = 4
    Qh = [torch.randn(B, D, generator=gq).to(torch.bfloat16).pin_memory() for _ in range(n_q)]
    Qd = [q.to(dev) for q in Qh]
    out_h = (torch.empty(B, k, dtype=torch.float32).pin_memory(), torch.empty(B, k, dtype=torch.int32).pin_memory())

    def step(q):
        return sharded.sharded_topk(q, E, k, id_offset=lo, mask_id=0)

    for it in range(max(args.warmup, 3)):
        step(Qd[it % n_q])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = _lib.launch_count()
    _lib.kernel_timer(["bdlru_fullsort_topk"])
    evs = []
    for it in range(args.steps):
        if E.numel() * 2 < (256 << 20):
            flush_l2(dev)  # shards below 2x L2 are flushed; larger ones evict themselves
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        step(Qd[it % n_q])
        e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    kt = _lib.kernel_timer_stop()["bdlru_fullsort_topk"]
    launches = _lib.launch_count() - n0
    if world > 1:
        dist.barrier()
    total_ms = sum(s.elapsed_time(e) for s, e in evs)
    tt = torch.tensor([total_ms], device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_per_step = float(tt) / args.steps

    # e2e: queries from pinned host memory, top-k lists back to the host, every step
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for it in range(args.steps):
        q = Qh[it % n_q].to(dev, non_blocking=True)
        s, i = step(q)
        out_h[0].copy_(s, non_blocking=True)
        out_h[1].copy_(i, non_blocking=True)
        torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te)
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        finish_distributed(world)
        return
    P = peaks()
    flops = 2.0 * B * (hi - lo) * D
    avg = sum(kt) / len(kt)
    tf = flops / (avg * 1e-3) / 1e12
    roofline = dict(bound="tensor", kernel="fullsort_kernel<UB=2,TOPK,K=10,NT=96,NSTG=2> (+ list merge)", achieved=tf,
                    peak=P["tf_sustained"], unit="TFLOP/s", frac=tf / P["tf_sustained"],
                    traffic=ncu_traffic(args.workload, "bdlru_fullsort_topk") if world == 1 else None,
                    peak_source=P["src"] + " (sustained bf16 cuBLAS; burst %.0f)" % P["tf"],
                    algorithmic_flops_per_launch=flops, avg_launch_ms=avg)
    base = None
    if not args.no_cpu and world == 1:   # cpu_baseline: rank 0 at N = 1 only
        base, _ = cpu_score_baseline(N, D, args.cpu_sample or (512 if N <= 1_000_000 else 64), k, steps=6, warmup=1)
    line = dict(metric="fullsort_scored_users_per_s", value=B / (ms_per_step * 1e-3), unit="users/s", n_gpus=world,
                steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms_per_step, higher_is_better=True,
                scaling="strong", vs_baseline=None, dtype="bf16", data="synthetic",
                config=dict(workload=f"{args.workload}: {B} users x {N} items, D={D}, top-{k}, column 0 masked, item "
                                     f"table row-sharded over {world} GPU(s)",
                            l2="shard > L2 evicts itself" if E.numel() * 2 >= (256 << 20) else "flushed between steps",
                            parallelism=f"item-shard x{world} + NCCL all-gather top-k merge" if world > 1 else "single"),
                e2e=dict(value=B * args.steps / e2e_s, unit="users/s", h2d_bytes_per_step=B * D * 2,
                         d2h_bytes_per_step=B * k * 8, ms_per_step=e2e_s / args.steps * 1e3),
                gpu_launches=launches, clocks=clocks, roofline=roofline, cpu_baseline=base)
    print(json.dumps(line), flush=True)
    finish_distributed(world)
