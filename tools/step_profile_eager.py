"""Kernel-level breakdown of an EAGER training step of a bench workload (torch.profiler / CUPTI).
    python tools/step_profile_eager.py strain1m"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from datamining_recblr_b200.recblr import RecBLR  # noqa: E402

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "strain1m"]
dev = torch.device("cuda")
torch.manual_seed(2020)
model = RecBLR(bench.make_config(w, dev), bench._DS(w["n_items"])).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True, capturable=True)
b = tuple(t.to(dev) for t in bench.synthetic_batch(w["B"], w["L"], w["n_items"], 1))
ex = {"item_id_list": b[0], "item_length": b[1], "item_id": b[2]}
model.train()


def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = model.calculate_loss(ex)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
N = 3
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"total device time per step: {tot / N / 1e3:.3f} ms over {sum(e.count for e in rows) / N:.0f} kernels")
for e in rows[:34]:
    print(f"{e.device_time_total / N / 1e3:8.3f} ms/step {e.count / N:5.1f}x  {e.key[:120]}")
