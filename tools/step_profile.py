"""Kernel-level time breakdown of the CUDA-graph training step of bench.py's `beauty` workload (torch.profiler / CUPTI,
warm, back-to-back replays): tells which kernels the 2 ms go to.   python tools/step_profile.py [workload]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from datamining_recblr_b200.recblr import RecBLR  # noqa: E402
from datamining_recblr_b200.train_step import GraphedTrainStep  # noqa: E402

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "beauty"]
dev = torch.device("cuda")
torch.manual_seed(2020)
with torch.device(dev):   # parameters are created (and initialised) on the GPU: a 10 M x 128 table takes seconds on the host
    model = RecBLR(bench.make_config(w, dev), bench._DS(w["n_items"]))
opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True, capturable=True)
b = tuple(t.to(dev) for t in bench.synthetic_batch(w["B"], w["L"], w["n_items"], 1))
ex = {"item_id_list": b[0], "item_length": b[1], "item_id": b[2]}
model.train()
if w.get("big"):   # the large-catalog step runs eagerly in bench.py too

    def step(inter):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = model.calculate_loss(inter)
        loss.backward()
        opt.step()
        return loss.detach()
else:
    step = GraphedTrainStep(model, opt, ex, autocast_dtype=torch.bfloat16)
for _ in range(5):
    step(ex)
torch.cuda.synchronize()
N = 3 if w.get("big") else 10
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        step(ex)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"total device time per step: {tot / N / 1e3:.3f} ms over {sum(e.count for e in rows) / N:.0f} kernels")
for e in rows[:45]:
    print(f"{e.device_time_total / N:9.1f} us/step {e.count / N:5.1f}x  {e.key[:110]}")
