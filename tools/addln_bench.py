"""Times ops.add_dropout_layernorm forward / backward at a training shape (rows x D, bf16 or fp32), reporting achieved
algorithmic GB/s (fwd: 3 tensors, bwd: 5 tensors of rows*D*esize).  BDLRU_LIB selects a variant library.

    python tools/addln_bench.py [rows] [D] [dtype] [p]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from datamining_recblr_b200 import ops  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8192 * 200
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    dt = torch.float32 if (len(sys.argv) > 3 and sys.argv[3] == "f32") else torch.bfloat16
    p = float(sys.argv[4]) if len(sys.argv) > 4 else 0.2
    dev = "cuda"
    x = torch.randn(rows, D, device=dev, dtype=dt).requires_grad_()
    r = torch.randn(rows, D, device=dev, dtype=dt).requires_grad_()
    g = torch.ones(D, device=dev, requires_grad=True)
    b = torch.zeros(D, device=dev, requires_grad=True)
    gy = torch.randn(rows, D, device=dev, dtype=dt)
    E = rows * D * x.element_size()

    def t(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        evs = []
        for _ in range(n):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            evs.append((s, e))
        torch.cuda.synchronize()
        ts = sorted(s.elapsed_time(e) for s, e in evs)
        return ts[len(ts) // 2]

    with torch.no_grad():
        tf = t(lambda: ops.add_dropout_layernorm(x, r, g, b, 1e-12, p, 7))
    y = ops.add_dropout_layernorm(x, r, g, b, 1e-12, p, 7)

    def bwd():
        x.grad = r.grad = g.grad = b.grad = None
        y.backward(gy, retain_graph=True)
    tb = t(bwd)
    print(f"{os.path.basename(os.environ.get('BDLRU_LIB', 'libbdlru.so')):32s} rows={rows} D={D} {str(dt)[6:]} p={p}: "
          f"fwd {tf:.3f} ms {3 * E / tf / 1e6:.0f} GB/s | bwd {tb:.3f} ms {5 * E / tb / 1e6:.0f} GB/s", flush=True)


if __name__ == "__main__":
    main()
