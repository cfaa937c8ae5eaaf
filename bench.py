#!/usr/bin/env python
"""bench.py — the hot path of BASELINE.json on N B200s of one node; prints ONE JSON line (rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload north_star|strain10m|strain1m|beauty|ml1m|score1m|score10m]

Workloads (BASELINE.json `configs`):
  north_star (default) = the largest single-GPU configuration, configs[4] "synthetic large-scale training": RecBLR with
           10 M items, L = 200, D = 128 (C = 256), 2 layers, batch 8 192 per GPU, bf16 autocast.  A step is RecBole's
           `_train_epoch` body for one batch: zero_grad -> calculate_loss (front end, 2 x (BD-LRU + FFN), full-softmax CE
           over all 10 M items) -> backward -> Adam.  The tied item table is ROW-SHARDED over the N ranks (fp32 master
           shard + optimizer state on the owner, replicated bf16 copy refreshed by an all-gather; sharded.ShardedItemTable),
           the layers are data parallel.  metric = BD-LRU fwd+bwd seq-tokens/s = N*B*L / step time (weak scaling).
           Sub-objects on the same line: `fullsort` = configs[3], 4 096 users scored against the SAME 10 M-row table,
           row-sharded over the N ranks with the NCCL list merge (users/s, strong scaling); `sweep` = configs[2] corners of
           the fused gate+scan fwd+bwd against the HBM roofline, fp32 and bf16 (N = 1 only); `vs_triton` = the reference's
           own Triton-scan layer timed on the same GPU (N = 1 only); `parity_check` = in-run assertion that the sharded
           path equals the single-table kernels; `cpu_baseline` (N = 1 only).
  strain10m / strain1m   the training line alone (1 M items = the 10x smaller probe).
  beauty   configs[1]: Amazon-Beauty shape — n_items 12 102, L = 50, D = 64, B = 2 048/GPU, table replicated, whole step
           as one CUDA graph, data-parallel gradient all-reduce.    ml1m: configs[0]'s shape through the same step.
  score1m / score10m   configs[3] alone: synthetic full-sort scoring, D = 128, 4 096 users, table row-sharded.

`--impl reference` times the CPU restatement of the reference (oracle/torch_port.py: literal pad, F.conv1d,
separate gate ops, sequential scan) with all host threads on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: n_items, L, D, layers, train B, eval B
    "beauty": dict(n_items=12102, L=50, D=64, layers=2, B=2048, eval_B=4096),
    "ml1m": dict(n_items=3417, L=200, D=64, layers=2, B=2048, eval_B=4096),
    # configs[4] (S-train): large catalog, data parallel + row-sharded full-softmax CE; strain1m is the 10x smaller probe
    "strain1m": dict(n_items=1_000_000, L=200, D=128, layers=2, B=8192, eval_B=4096, big=True),
    "strain10m": dict(n_items=10_000_000, L=200, D=128, layers=2, B=8192, eval_B=4096, big=True),
}
SCORE_WORKLOADS = {"score1m": 1_000_000, "score10m": 10_000_000}


# ----------------------------------------------------------------------------- helpers
def peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    except Exception:
        return dict(hbm=6650.0, tf=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.t0 = index, None, [], 0.0

    def mark(self):
        """Samples taken before this call are dropped.  start() is called BEFORE the warm-up steps (the fork of this
        process and nvidia-smi's NVML start-up stalled kernel launches for ~20 ms, which landed in the first timed step
        when it was started right in front of the timed region), mark() right in front of the timed region."""
        self.t0 = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            if ts < self.t0:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def finish_distributed(world):
    """Collective teardown that cannot hang the launcher: every rank syncs, meets at a barrier (ranks != 0 wait here for
    rank 0's rank-0-only legs), then destroys the process group; if NCCL teardown stalls (seen with captured graphs
    still holding communicator resources) a daemon timer ends the process — the JSON line is already printed."""
    if world <= 1:
        return
    import torch.distributed as dist
    sys.stdout.flush()
    torch.cuda.synchronize()
    try:
        dist.barrier()
    except Exception:
        pass
    t = threading.Timer(20.0, lambda: os._exit(0))
    t.daemon = True
    t.start()
    dist.destroy_process_group()
    t.cancel()


def ncu_traffic(workload, kernel):
    """dram bytes per launch of `kernel` from the committed ncu capture of this workload (None if there is none)."""
    for name in ("r2_dram_traffic.json", "r1_dram_traffic.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))[workload][kernel]
        except Exception:
            continue
    return None


def refuse_tuning_environment():
    """A timed kernel must not be steerable from the environment: refuse to run with any BDLRU_* variable set or with a
    -DBDLRU_TUNING build of the library (the only kind that reads them)."""
    bad = sorted(k for k in os.environ if k.startswith("BDLRU_"))
    if bad:
        raise SystemExit(f"bench.py refuses to run with tuning variables set: {bad}")


def bench_header():
    """Which binary produced the numbers: ABI version, source digest baked into libbdlru.so, and whether that digest
    matches the sources lying next to it."""
    from datamining_recblr_b200 import _lib
    info = _lib.build_info()
    if info["tuning"]:
        raise SystemExit("bench.py refuses to time a -DBDLRU_TUNING build of libbdlru.so")
    return {"library": info}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def make_config(w, dev, dropout=0.2):
    cfg = dict(hidden_size=w["D"], num_layers=w["layers"], dropout_prob=dropout, expand=2, d_conv=4, loss_type="CE",
               USER_ID_FIELD="user_id", ITEM_ID_FIELD="item_id", LIST_SUFFIX="_list",
               ITEM_LIST_LENGTH_FIELD="item_length", NEG_PREFIX="neg_", MAX_ITEM_LIST_LENGTH=w["L"], device=dev)

    class Cfg(dict):  # RecBole's Config returns None for unknown keys
        def __getitem__(self, k):
            return self.get(k, None)
    return Cfg(cfg)


class _DS:
    def __init__(self, n):
        self.n = n

    def num(self, field):
        return self.n


def synthetic_batch(B, L, n_items, seed):
    """ids uniform in [1, n_items), lengths uniform in [min(5, L), L], right-padded with 0, targets uniform (§8d)."""
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(min(5, L), L + 1, (B,), generator=g)
    seq = torch.randint(1, n_items, (B, L), generator=g)
    seq = seq * (torch.arange(L)[None, :] < lens[:, None])
    pos = torch.randint(1, n_items, (B,), generator=g)
    return seq, lens, pos


# ----------------------------------------------------------------------------- reference arm (CPU)
def cpu_train_baseline(w, steps, warmup, sample_B):
    """The reference's PyTorch path with the sequential scan on the host cores (oracle/torch_port.py), same step
    (zero_grad, CE loss over all items, backward, Adam), on a bounded sample of sample_B sequences per step."""
    from oracle import torch_port as TP
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wts = TP.init_weights(w["n_items"], w["D"], w["layers"], seed=2020)
    params = [v.requires_grad_(True) for v in wts.values()]
    opt = torch.optim.Adam(params, lr=1e-3)
    seq, lens, pos = synthetic_batch(sample_B, w["L"], w["n_items"], seed=2020)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = TP.ce_loss(wts, seq, lens, pos, w["layers"], dropout_p=0.2)
        loss.backward()
        opt.step()
        loss.item()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    return dict(value=sample_B * w["L"] / t, unit="seq-tokens/s", cores=cores, kind="port",
                sample=f"{sample_B} sequences x L={w['L']} per step ({len(times)} timed steps, {t:.2f} s/step), "
                       f"fp32, torch {torch.__version__} CPU, oracle/torch_port.py"), t


def cpu_score_baseline(n_rows, D, users, k, steps, warmup):
    """Reference full-sort eval on CPU: dense fp32 scores = Q @ E^T, scores[:, 0] = -inf, torch.topk (RecBLR.py:118-122
    + RecBole's collector) on a bounded number of users against the full table."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(2020)
    E = torch.randn(n_rows, D, generator=g) * 0.02
    Q = torch.randn(users, D, generator=g)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        s = Q @ E.T
        s[:, 0] = float("-inf")
        torch.topk(s, k)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    return dict(value=users / t, unit="users/s", cores=cores, kind="port",
                sample=f"{users} users x {n_rows} items per step ({len(times)} timed steps, {t:.2f} s/step), fp32 "
                       f"matmul + torch.topk on CPU"), t


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    wname = "strain10m" if args.workload == "north_star" else args.workload
    if wname in WORKLOADS:
        w = WORKLOADS[wname]
        # bounded: each CPU step costs seconds (10 M-row table: CE over all items, dense dE, Adam), so at most 5 timed steps
        args.steps = min(args.steps, 5 if w.get("big") else 20)
        base, t = cpu_train_baseline(w, args.steps, 1 if w.get("big") else max(1, min(args.warmup, 2)),
                                     sample_B=args.cpu_sample or (32 if w.get("big") else 1024))
        line = dict(metric="bdlru_fwd_bwd_seq_tokens_per_s", value=base["value"], unit="seq-tokens/s",
                    config=dict(workload=f"{wname}: RecBLR n_items={w['n_items']} L={w['L']} D={w['D']} C={2 * w['D']} "
                                         f"layers={w['layers']} batch {w['B']}/GPU, train step = zero_grad + calculate_loss "
                                         f"(CE over all items) + backward + Adam",
                                sample=base["sample"]),
                    dtype="f32")
    else:
        n = SCORE_WORKLOADS[args.workload]
        base, t = cpu_score_baseline(n, 128, args.cpu_sample or 64, 10, args.steps, max(1, min(args.warmup, 2)))
        line = dict(metric="fullsort_scored_users_per_s", value=base["value"], unit="users/s",
                    config=dict(workload=f"{args.workload}: {n} items D=128 top-10"), dtype="f32")
    line.update(impl="reference", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=t * 1e3,
                higher_is_better=True, scaling="weak", vs_baseline=None, data="synthetic", cpu_baseline=base,
                e2e=dict(value=base["value"], unit=base["unit"], h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    emit(line)


# ----------------------------------------------------------------------------- our arm: training step
def parity_check_sharded_ce(sit, dev, n=512):
    """ASSERTS, before any timing, that the row-sharded CE statistics (per-shard kernel + NCCL combine) equal the
    single-table kernel run on the replicated copy: global LSE and positive logit of n seeded users."""
    from datamining_recblr_b200 import ops, sharded
    g = torch.Generator().manual_seed(77)
    q = torch.randn(n, sit.D, generator=g).to(torch.bfloat16).to(dev)
    pos = torch.randint(1, sit.n_items, (n,), generator=g).to(dev)
    m, s, pl = ops.fullsort_ce_stats(q, sit.shard_bf16(), pos, id_offset=sit.lo)
    lse, pl = sharded.combine_ce_stats(m, s, pl, sit.group)
    m1, s1, pl1 = ops.fullsort_ce_stats(q, sit.table_bf16[:sit.n_items], pos)
    lse1 = m1 + torch.log(s1)
    e_lse = float(((lse - lse1).abs() / lse1.abs().clamp_min(1e-6)).max())
    e_pl = float((pl - pl1).abs().max() / pl1.abs().max().clamp_min(1e-6))
    assert e_lse <= 1e-6 and e_pl <= 1e-6, f"parity_check: sharded CE differs from the single table (lse {e_lse}, pos {e_pl})"
    return {"ce_lse_max_rel": e_lse, "ce_pos_logit_max_rel": e_pl, "ce_users": n}


def run_train(args):
    import torch.distributed as dist
    from datamining_recblr_b200 import _lib, sharded
    from datamining_recblr_b200.recblr import RecBLR
    from datamining_recblr_b200.timing import flush_l2

    rank, world, local = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    north = args.workload == "north_star"
    wname = "strain10m" if north else args.workload
    w = WORKLOADS[wname]
    B, L, D = w["B"], w["L"], w["D"]
    torch.manual_seed(2020)
    big = bool(w.get("big"))
    cfg = make_config(w, dev)
    with torch.device(dev):   # parameters are created on the GPU (same seed => identical on every rank)
        model = RecBLR(cfg, _DS(w["n_items"]))
    sit = None
    if big:   # tied item table row-sharded over the ranks (world 1: same code, no collectives)
        sit = sharded.shard_item_table(model)
        torch.cuda.empty_cache()
    if sit is not None:   # dense parameters: torch's fused Adam; table shard: Adam + bf16 refresh in one kernel + all-gather
        opt = sharded.ShardedTableOptimizer(model, lr=1e-3)
    else:
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True, capturable=True)
    dense = sharded.dense_parameters(model)
    amp = args.dtype == "bf16"
    assert amp or sit is None, "the sharded-table workloads run in bf16"

    # distinct batches per step (and per rank)
    n_batches = 2 if big else 4
    host = [synthetic_batch(B, L, w["n_items"], seed=2020 + 97 * rank + i) for i in range(n_batches)]
    host = [tuple(t.pin_memory() for t in b) for b in host]
    devb = [tuple(t.to(dev) for t in b) for b in host]

    def allreduce_grads(ps):
        if sit is not None:   # loss is the GLOBAL mean: dense gradients are summed; the table gradient is owner-local
            sharded.allreduce_gradients(dense, average=False)
        else:
            sharded.allreduce_gradients(ps, average=True)   # one flat NCCL call, grads become views

    def eager_step(batch):
        inter = {"item_id_list": batch[0], "item_length": batch[1], "item_id": batch[2]}
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            loss = model.calculate_loss(inter)
        loss.backward()
        if world > 1:
            allreduce_grads([p for p in model.parameters()])
        opt.step()
        return loss.detach()

    parity = None
    if sit is not None:
        parity = parity_check_sharded_ce(sit, dev)

    model.train()
    graphed = None
    if not args.no_graph and not big:   # the large-catalog step is not launch-bound; it runs eagerly
        from datamining_recblr_b200.train_step import GraphedTrainStep
        ex = {"item_id_list": devb[0][0], "item_length": devb[0][1], "item_id": devb[0][2]}
        graphed = GraphedTrainStep(model, opt, ex, autocast_dtype=torch.bfloat16 if amp else None,
                                   grad_hook=allreduce_grads if world > 1 else None)

    graphed_launches = 0
    if graphed is not None:  # launches of libbdlru kernels in one step (same code path as the captured graph)
        c0 = _lib.launch_count()
        eager_step(devb[0])
        graphed_launches = _lib.launch_count() - c0

    def step(batch):
        if graphed is None:
            return eager_step(batch)
        return graphed({"item_id_list": batch[0], "item_length": batch[1], "item_id": batch[2]})

    model.train()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(max(args.warmup, 3)):
        flush_l2(dev)   # as in the timed loop: the flush buffer is allocated here, not inside the first timed step
        step(devb[i % n_batches])
    torch.cuda.synchronize()

    # ---- timed region: K steps, inputs resident in HBM, CUDA events per step, L2 flushed between steps
    timed = ["bdlru_gated_scan_fwd", "bdlru_gated_scan_bwd", "bdlru_conv1d_fwd", "bdlru_conv1d_bwd",
             "bdlru_embed_ln_fwd", "bdlru_embed_ln_bwd", "bdlru_embed_ln_bwd_rows", "bdlru_scatter_add_rows",
             "bdlru_fullsort_ce_fwd", "bdlru_fullsort_ce_bwd", "bdlru_fullsort_rowmax", "bdlru_fullsort_ce_fwd_dq",
             "bdlru_add_ln_fwd", "bdlru_add_ln_bwd", "bdlru_colsum",
             "bdlru_silu_dropout_fwd", "bdlru_silu_dropout_bwd"]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    # per-kernel launch durations: event pairs around the C-ABI calls.  Events cannot be recorded inside a graph
    # replay, so with the graphed step they are taken from eager steps of the same work right before the timed region.
    ktimes = None
    if graphed is not None:
        _lib.kernel_timer(timed)
        for i in range(3):
            flush_l2(dev)
            eager_step(devb[i % n_batches])
        ktimes = _lib.kernel_timer_stop()
        eager_steps = 3
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
    n0 = _lib.launch_count()
    if graphed is None:
        _lib.kernel_timer(timed)
        eager_steps = args.steps
    sampler.mark()
    evs = []
    for i in range(args.steps):
        flush_l2(dev)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        step(devb[i % n_batches])
        e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    if graphed is None:
        ktimes = _lib.kernel_timer_stop()
    launches = _lib.launch_count() - n0
    if graphed is not None:  # replays do not pass through the host launch counter: count what the graph contains
        launches = graphed_launches * args.steps
    if world > 1:
        dist.barrier()
    per_step = [s.elapsed_time(e) for s, e in evs]
    tt = torch.tensor([sum(per_step), statistics.median(per_step), min(per_step)], device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, med_ms, min_ms = (float(x) for x in tt)
    ms_per_step = total_ms / args.steps
    value = world * B * L / (ms_per_step * 1e-3)

    # ---- e2e: same step through the public API from pinned HOST buffers, loss read back every step
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        hb = host[i % n_batches]
        db = tuple(t.to(dev, non_blocking=True) for t in hb)
        step(db).item()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te)
    h2d = sum(t.numel() * t.element_size() for t in host[0])

    # ---- the other half of the metric: full-sort scored users/s
    fullsort = None
    if sit is not None:
        # configs[3] on the table just trained: 4 096 users against all n_items rows, row-sharded over the ranks
        from bench_score import score_leg
        fullsort = score_leg(sit.shard_bf16(), sit.lo, sit.n_items, steps=20 if north else 10, warmup=3,
                             full_table=sit.table_bf16[:sit.n_items])
    elif rank == 0:
        fullsort = model_fullsort_leg(model, w, dev, amp, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        del graphed
        finish_distributed(world)
        return

    # ---- roofline of the dominant kernel of libbdlru.so in the step (live CUDA-event durations)
    P = peaks()
    es = 2 if amp else 4
    C = 2 * D
    E_l = B * L * C  # elements of one [B, L, C] activation
    # algorithmic bytes per launch: z-fused gated scan fwd reads x', r, i, z and writes h, y;
    # bwd reads x', r, i, z, h, g and writes dx', dr, di, dz
    tab_es = 2 if sit is not None else 4
    alg = {"bdlru_gated_scan_fwd": 6 * E_l * es, "bdlru_gated_scan_bwd": 10 * E_l * es,
           "bdlru_conv1d_fwd": 2 * E_l * es, "bdlru_conv1d_bwd": 4 * E_l * es,
           "bdlru_embed_ln_fwd": B * L * (8 + D * tab_es + D * es), "bdlru_embed_ln_bwd": B * L * (8 + D * tab_es + D * es + D * 4),
           "bdlru_embed_ln_bwd_rows": B * L * (8 + D * tab_es + 2 * D * es),
           "bdlru_add_ln_fwd": 3 * B * L * D * es, "bdlru_add_ln_bwd": 5 * B * L * D * es,
           "bdlru_silu_dropout_fwd": 2 * B * L * 4 * D * es, "bdlru_silu_dropout_bwd": 3 * B * L * 4 * D * es}
    per_kernel = {}
    for name, ts in ktimes.items():
        if ts:
            per_kernel[name] = dict(calls_per_step=len(ts) / eager_steps, avg_ms=sum(ts) / len(ts),
                                    share_of_step=(sum(ts) / eager_steps) / ms_per_step,
                                    gbs=(alg[name] / (sum(ts) / len(ts)) / 1e6) if name in alg else None)
            if name in alg:
                per_kernel[name]["hbm_frac"] = per_kernel[name]["gbs"] / P["hbm"]
    Bce = B * world if sit is not None else B
    Nce = sit.n_local if sit is not None else w["n_items"]
    # CREDITED flops follow SURVEY §8d (a materialising implementation: logits GEMM 2*B*N*D, dQ 2*B*N*D, dE 2*B*N*D);
    # EXECUTED flops are what the non-materialising kernels run.  With the fused forward (ce_fwd_dq: logits + P.E in one
    # pass) the backward entry point runs the dE pass only, which must recompute the logits (executed 4, credited 2).
    BND = float(Bce) * Nce * D
    fused_fwd = "bdlru_fullsort_ce_fwd_dq" in per_kernel
    flops = {"bdlru_fullsort_ce_fwd": 2.0 * BND, "bdlru_fullsort_ce_fwd_dq": 4.0 * BND,
             "bdlru_fullsort_ce_bwd": (2.0 if fused_fwd else 4.0) * BND}
    executed = {"bdlru_fullsort_ce_fwd": 2.0 * BND, "bdlru_fullsort_ce_fwd_dq": 4.0 * BND,
                "bdlru_fullsort_ce_bwd": (4.0 if fused_fwd else 8.0) * BND,
                "bdlru_fullsort_rowmax": 2.0 * BND / 16}   # sampled reference maximum: every 16th tile, nothing credited
    for name, f in executed.items():
        if name in per_kernel:
            sec = per_kernel[name]["avg_ms"] * 1e-3
            per_kernel[name]["tflops"] = flops.get(name, 0.0) / sec / 1e12
            per_kernel[name]["tensor_frac"] = per_kernel[name]["tflops"] / P["tf_sustained"]
            per_kernel[name]["tflops_executed"] = f / sec / 1e12
    cand = [n for n in per_kernel if n in alg or n in flops]
    dom = max(cand, key=lambda n: per_kernel[n]["avg_ms"] * per_kernel[n]["calls_per_step"])
    if dom in flops:
        roofline = dict(bound="tensor", kernel=dom, achieved=per_kernel[dom]["tflops"], peak=P["tf_sustained"],
                        unit="TFLOP/s", frac=per_kernel[dom]["tflops"] / P["tf_sustained"],
                        traffic=ncu_traffic(wname, dom),
                        peak_source=P["src"] + " (sustained bf16 cuBLAS: the kernel is timed inside a long step)",
                        algorithmic_flops_per_launch=flops[dom],
                        avg_launch_ms=per_kernel[dom]["avg_ms"],
                        executed_frac=per_kernel[dom].get("tflops_executed", per_kernel[dom]["tflops"]) / P["tf_sustained"],
                        note="credited flops (SURVEY §8d) count a materialising implementation: 2*B*N*D each for the logits, dQ "
                             "and dE GEMMs.  The fused forward (bdlru_fullsort_ce_fwd_dq) runs logits + dQ in one pass (4 "
                             "credited = 4 executed); the backward entry point then runs the dE pass only, which has to "
                             "recompute the logits: 2 credited, 4 executed (executed_frac)")
    else:
        roofline = dict(bound="hbm", kernel=dom, achieved=per_kernel[dom]["gbs"], peak=P["hbm"], unit="GB/s",
                        frac=per_kernel[dom]["gbs"] / P["hbm"], traffic=ncu_traffic(wname, dom),
                        peak_source=P["src"],
                        algorithmic_bytes_per_launch=alg[dom], avg_launch_ms=per_kernel[dom]["avg_ms"],
                        note="events bracket the C-ABI call on the launching stream (includes its dLambda/dh0 reduction "
                             "launch)" + ("" if big else "; working set is L2-resident at this shape, see DESIGN.md"))

    sweep = vs_triton = None
    if north and world == 1:
        import gc
        del opt
        if sit is not None:
            sit._grad = None
        gc.collect()
        torch.cuda.empty_cache()
        from bench_legs import sweep_leg, triton_leg
        sweep = sweep_leg(P["hbm"])
        vs_triton = triton_leg()
    base, _ = (cpu_train_baseline(w, steps=2 if big else 6, warmup=1, sample_B=args.cpu_sample or (32 if big else 1024))
               if not args.no_cpu and world == 1 else (None, 0))   # cpu_baseline: rank 0 at N = 1 only
    par = (f"dp{world} layers + item table row-sharded x{world} (fp32 master + Adam state on the owner, replicated bf16 "
           f"copy all-gathered per step; CE dE owner-local, dQ reduce-scatter, embedding rows all-gather + owner scatter)"
           if sit is not None and world > 1 else (f"dp{world}" if world > 1 else "single"))
    line = dict(metric="bdlru_fwd_bwd_seq_tokens_per_s", value=value, unit="seq-tokens/s", n_gpus=world,
                steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms_per_step, ms_median=med_ms, ms_min=min_ms,
                step_ms=[round(x, 2) for x in per_step],
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16" if amp else "f32", data="synthetic",
                config=dict(workload=f"{wname}: RecBLR n_items={w['n_items']} L={L} D={D} C={C} "
                                     f"layers={w['layers']} batch {B}/GPU, train step = zero_grad + calculate_loss "
                                     f"(CE over all items) + backward + Adam",
                            l2="flushed between steps (256 MB write)" + ("" if big else "; per-step working set < L2"),
                            parallelism=par, ce_impl="sharded-table" if sit is not None else model.ce_impl,
                            scan_state="fp32", table="fp32 master + bf16 compute copy" if sit is not None else "fp32",
                            launch="CUDA graph replay of the whole step" if graphed is not None else "eager"),
                e2e=dict(value=world * B * L * args.steps / e2e_s, unit="seq-tokens/s", h2d_bytes_per_step=h2d,
                         d2h_bytes_per_step=4, ms_per_step=e2e_s / args.steps * 1e3),
                gpu_launches=launches, clocks=clocks, roofline=roofline, kernels=per_kernel, fullsort=fullsort,
                parity_check=({**(fullsort or {}).get("parity_check", {}), **parity, "status": "ok"}
                              if parity is not None else None),
                sweep=sweep, vs_triton=vs_triton, cpu_baseline=base, **bench_header())
    emit(line)
    del graphed
    finish_distributed(world)


def model_fullsort_leg(model, w, dev, amp, steps):
    """Small-catalog workloads: users/s of model forward + fused scoring + top-10 at the eval batch (rank 0)."""
    from datamining_recblr_b200 import ops
    from datamining_recblr_b200.timing import flush_l2
    model.eval()
    EB, L, D = w["eval_B"], w["L"], w["D"]
    eb = tuple(t.to(dev) for t in synthetic_batch(EB, L, w["n_items"], seed=7))
    inter = {"item_id_list": eb[0], "item_length": eb[1], "item_id": eb[2]}

    def score():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            return model.full_sort_topk(inter, 10)
    assert ops.fullsort_supported(D)
    for _ in range(3):
        score()
    ts = []
    for _ in range(max(steps, 5)):
        flush_l2(dev)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        score()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    model.train()
    return dict(metric="fullsort_scored_users_per_s", value=EB / (statistics.median(ts) * 1e-3), unit="users/s",
                users=EB, n_items=w["n_items"], k=10, ms=statistics.median(ts), ms_min=min(ts), fused_topk=True,
                includes="model forward + fused scoring + top-10")


def reserve_stdout():
    """The contract is ONE JSON line on stdout: everything else that may write to file descriptor 1 during the run (NCCL's
    version banner, library chatter) is sent to stderr; emit() writes the line to the real stdout.  The saved descriptor
    lives on `sys` because this file is loaded twice (as __main__ and, from bench_score.py, as module `bench`)."""
    if getattr(sys, "_bdlru_json_fd", None) is None:
        sys.stdout.flush()
        sys._bdlru_json_fd = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    fd = getattr(sys, "_bdlru_json_fd", None)
    os.write(1 if fd is None else fd, (json.dumps(line) + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed steps (default: 10 large-catalog / 200 small / 50 scoring)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="north_star", choices=["north_star"] + list(WORKLOADS) + list(SCORE_WORKLOADS))
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="sequences (train) / users (scoring) per CPU step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="run the training step eagerly instead of as a CUDA graph")
    args = ap.parse_args()
    reserve_stdout()
    train_name = "strain10m" if args.workload == "north_star" else args.workload
    if args.steps <= 0:
        big_w = train_name in WORKLOADS and WORKLOADS[train_name].get("big")
        args.steps = (3 if args.impl == "reference" and big_w else 5 if args.impl == "reference"
                      else (10 if big_w else (200 if train_name in WORKLOADS else 50)))
    if args.impl == "reference":
        return run_reference(args)
    refuse_tuning_environment()
    if train_name in WORKLOADS:
        return run_train(args)
    from bench_score import run_score  # noqa: WPS433  (kept separate: needs the tcgen05 kernels)
    return run_score(args)


if __name__ == "__main__":
    main()
