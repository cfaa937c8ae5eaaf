"""Kernel-level time breakdown of bench.py's training step (torch.profiler / CUPTI, warm, back-to-back steps): tells which
kernels the step time goes to, NCCL collectives included.

    python tools/step_profile.py [workload]                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node N tools/step_profile.py strain10m    # N GPUs (rank 0 prints)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from datamining_recblr_b200 import sharded  # noqa: E402
from datamining_recblr_b200.recblr import RecBLR  # noqa: E402
from datamining_recblr_b200.train_step import GraphedTrainStep  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "beauty"
w = bench.WORKLOADS[name]
rank, world, local = bench.dist_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=dev)
torch.manual_seed(2020)
with torch.device(dev):   # parameters are created (and initialised) on the GPU: a 10 M x 128 table takes seconds on the host
    model = RecBLR(bench.make_config(w, dev), bench._DS(w["n_items"]))
big = bool(w.get("big"))
sit = sharded.shard_item_table(model) if big else None
if sit is not None:
    opt = sharded.ShardedTableOptimizer(model, lr=1e-3)
else:
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True, capturable=True)
dense = sharded.dense_parameters(model)
b = tuple(t.to(dev) for t in bench.synthetic_batch(w["B"], w["L"], w["n_items"], 1 + rank))
ex = {"item_id_list": b[0], "item_length": b[1], "item_id": b[2]}
model.train()
if big:   # the large-catalog step runs eagerly in bench.py too

    def step(inter):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = model.calculate_loss(inter)
        loss.backward()
        if world > 1:
            sharded.allreduce_gradients(dense, average=False)
        opt.step()
        return loss.detach()
else:
    step = GraphedTrainStep(model, opt, ex, autocast_dtype=torch.bfloat16)
for _ in range(4):
    step(ex)
torch.cuda.synchronize()
N = 3 if big else 10
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        step(ex)
    torch.cuda.synchronize()
if rank == 0:
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
    tot = sum(e.device_time_total for e in rows)
    print(f"[{name}, world {world}] total device time per step: {tot / N / 1e3:.3f} ms over "
          f"{sum(e.count for e in rows) / N:.0f} kernels")
    for e in rows[:45]:
        print(f"{e.device_time_total / N:9.1f} us/step {e.count / N:5.1f}x  {e.key[:110]}")
if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
