"""Full-sort scoring legs of bench.py (BASELINE.json configs[3]): 4 096 users x N items, D = 128, bf16 operands / fp32
accumulate, top-10 with RecBole's column-0 mask.  The item table is row-sharded over the ranks (rank g owns rows
[g*N/G, (g+1)*N/G)); every rank scores all users against its shard with the fused tcgen05 GEMM + streaming top-k kernel,
the packed (scores | ids) lists are exchanged with ONE NCCL all-gather and merged per user by (score desc, id asc), read in
place — identical to the 1-GPU result by construction.  A step scores the 4 096 users once; value = users /
step time (fixed catalog => strong scaling in N).  At N > 1 the whole step (kernel + all-gather + merge) is replayed as one
CUDA graph.

  score_leg(...)   the measurement, on a shard the caller provides (bench.py's default line passes the trained table's shard)
  run_score(args)  bench.py --workload score1m|score10m
"""
import json
import statistics
import time

import torch


def parity_check_scoring(q, shard, lo, k, full_table):
    """ASSERTS, before any timing, that the sharded path (per-shard kernel + NCCL exchange + merge) returns exactly the
    ids/scores of the single-table kernel on a slice of users.  `full_table`: the whole bf16 table on this rank (the
    replicated copy), or None when the rank only holds its shard (then rank-consistency of the merged lists is checked)."""
    import torch.distributed as dist
    from datamining_recblr_b200 import ops, sharded
    n = min(512, q.shape[0])
    s_sh, i_sh = sharded.sharded_topk(q[:n].contiguous(), shard, k, id_offset=lo, mask_id=0)
    out = {"users": n}
    if full_table is not None:
        s_1, i_1 = ops.fullsort_topk(q[:n].contiguous(), full_table, k, mask_id=0)
        assert torch.equal(i_sh, i_1), "parity_check: sharded top-k ids differ from the single-table kernel"
        assert torch.equal(s_sh, s_1), "parity_check: sharded top-k scores differ from the single-table kernel"
        out["topk_vs_single_table"] = "identical"
    if dist.is_initialized() and dist.get_world_size() > 1:
        ref = i_sh.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(ref, i_sh), "parity_check: merged lists differ between ranks"
        out["topk_rank_consistency"] = "identical"
    return out


def score_leg(E, lo, n_total, steps, warmup, k=10, B=4096, full_table=None, graph=None):
    """E: this rank's bf16 [rows, D] shard whose row 0 has global id `lo`.  Returns the rank-0 result dict (None on
    other ranks); every rank must call it."""
    import torch.distributed as dist
    from datamining_recblr_b200 import _lib, sharded
    from datamining_recblr_b200.timing import flush_l2
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    dev = E.device
    D = E.shape[1]
    gq = torch.Generator().manual_seed(2020)
    n_q = 4
    Qh = [torch.randn(B, D, generator=gq).to(torch.bfloat16).pin_memory() for _ in range(n_q)]
    Qd = [q.to(dev) for q in Qh]
    out_h = (torch.empty(B, k, dtype=torch.float32).pin_memory(), torch.empty(B, k, dtype=torch.int32).pin_memory())
    parity = parity_check_scoring(Qd[0], E, lo, k, full_table)

    def eager(q):
        return sharded.sharded_topk(q, E, k, id_offset=lo, mask_id=0)

    use_graph = (world > 1) if graph is None else graph
    q_static = Qd[0].clone()
    if use_graph:   # kernel + all-gather + merge captured once: one launch per step
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                eager(q_static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            res = eager(q_static)

        def step(q):
            q_static.copy_(q, non_blocking=True)
            g.replay()
            return res
    else:
        g = None
        step = eager

    for it in range(max(warmup, 3)):
        step(Qd[it % n_q])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n0 = _lib.launch_count()
    c0 = _lib.launch_count()
    eager(Qd[0])
    per_step_launches = _lib.launch_count() - c0
    # kernel-only duration (events around the C-ABI call, eager, on the launching stream)
    _lib.kernel_timer(["bdlru_fullsort_topk"])
    for it in range(5):
        if E.numel() * 2 < (256 << 20):
            flush_l2(dev)
        eager(Qd[it % n_q])
    kt = _lib.kernel_timer_stop()["bdlru_fullsort_topk"]
    if world > 1:
        dist.barrier()
    evs = []
    for it in range(steps):
        if E.numel() * 2 < (256 << 20):
            flush_l2(dev)  # shards below 2x L2 are flushed; larger ones evict themselves
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        step(Qd[it % n_q])
        e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ts = [s.elapsed_time(e) for s, e in evs]
    tt = torch.tensor([sum(ts), statistics.median(ts), min(ts)], device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, med_ms, min_ms = (float(x) for x in tt)
    ms_per_step = total_ms / steps

    # e2e: queries from pinned host memory, top-k lists back to the host, every step
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    q_in = torch.empty_like(Qd[0])   # the step's device input buffer, filled from pinned host memory every step
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for it in range(steps):
        q_in.copy_(Qh[it % n_q], non_blocking=True)
        s, i = step(q_in)
        out_h[0].copy_(s, non_blocking=True)
        out_h[1].copy_(i, non_blocking=True)
        torch.cuda.synchronize()
    ev1.record()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_dev_ms = ev0.elapsed_time(ev1) / steps   # the same loop on the device clock (host stalls show as the difference)
    te = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te)
    del g
    if rank != 0:
        return None
    flops = 2.0 * B * E.shape[0] * D
    avg = sum(kt) / len(kt)
    return dict(metric="fullsort_scored_users_per_s", value=B / (ms_per_step * 1e-3), unit="users/s", users=B,
                n_items=n_total, k=k, ms_per_step=ms_per_step, ms_median=med_ms, ms_min=min_ms, steps=steps,
                scaling="strong", launch="CUDA graph replay (kernel + all-gather + merge)" if use_graph else "eager",
                e2e=dict(value=B * steps / e2e_s, unit="users/s", h2d_bytes_per_step=B * D * 2,
                         d2h_bytes_per_step=B * k * 8, ms_per_step=e2e_s / steps * 1e3, device_ms_per_step=e2e_dev_ms),
                kernel=dict(name="fullsort_kernel<UB=2,TOPK,K=10,NT=96,NSTG=2> + list merge", avg_launch_ms=avg,
                            rows_per_rank=E.shape[0], algorithmic_flops_per_launch=flops,
                            tflops=flops / (avg * 1e-3) / 1e12),
                gpu_launches=per_step_launches * steps, parity_check=dict(status="ok", **parity))


def run_score(args):
    import torch.distributed as dist
    from bench import SCORE_WORKLOADS, ClockSampler, bench_header, cpu_score_baseline, dist_env, emit, \
        finish_distributed, ncu_traffic, peaks

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
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    leg = score_leg(E, lo, N, args.steps, args.warmup, k=k, B=B)
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        finish_distributed(world)
        return
    P = peaks()
    tf = leg["kernel"]["tflops"]
    roofline = dict(bound="tensor", kernel=leg["kernel"]["name"], achieved=tf, peak=P["tf_sustained"], unit="TFLOP/s",
                    frac=tf / P["tf_sustained"], frac_of_burst=tf / P["tf"],
                    traffic=ncu_traffic(args.workload, "bdlru_fullsort_topk") if world == 1 else None,
                    peak_source=P["src"] + " (sustained bf16 cuBLAS; burst %.0f)" % P["tf"],
                    algorithmic_flops_per_launch=leg["kernel"]["algorithmic_flops_per_launch"],
                    avg_launch_ms=leg["kernel"]["avg_launch_ms"])
    base = None
    if not args.no_cpu and world == 1:   # cpu_baseline: rank 0 at N = 1 only
        base, _ = cpu_score_baseline(N, D, args.cpu_sample or (512 if N <= 1_000_000 else 64), k, steps=6, warmup=1)
    line = dict(metric="fullsort_scored_users_per_s", value=leg["value"], unit="users/s", n_gpus=world,
                steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=leg["ms_per_step"], ms_median=leg["ms_median"],
                ms_min=leg["ms_min"], higher_is_better=True, scaling="strong", vs_baseline=None, dtype="bf16",
                data="synthetic",
                config=dict(workload=f"{args.workload}: {B} users x {N} items, D={D}, top-{k}, column 0 masked, item "
                                     f"table row-sharded over {world} GPU(s)",
                            l2="shard > L2 evicts itself" if E.numel() * 2 >= (256 << 20) else "flushed between steps",
                            parallelism=f"item-shard x{world} + one NCCL all-gather of packed lists + in-place merge"
                            if world > 1 else "single", launch=leg["launch"]),
                e2e=leg["e2e"], gpu_launches=leg["gpu_launches"], clocks=clocks, roofline=roofline,
                parity_check=leg["parity_check"], cpu_baseline=base, **bench_header())
    emit(line)
    finish_distributed(world)
