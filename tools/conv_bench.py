"""Times the channel-last causal conv1d (+bias, +SiLU) forward / backward at a training shape on the x half of a packed
[B, T, 2C] projection (row stride 2C), reporting achieved algorithmic GB/s (fwd 2 units, bwd 3 units of B*T*C*esize).

    python tools/conv_bench.py [B] [T] [C] [dtype]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from datamining_recblr_b200 import ops  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    C = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    dt = torch.float32 if (len(sys.argv) > 4 and sys.argv[4] == "f32") else torch.bfloat16
    dev = "cuda"
    xz = torch.randn(B, T, 2 * C, device=dev, dtype=dt)
    x = xz[..., :C].detach().requires_grad_()
    w = torch.randn(C, 4, device=dev, requires_grad=True)
    b = torch.randn(C, device=dev, requires_grad=True)
    gy = torch.randn(B, T, C, device=dev, dtype=dt)
    E = B * T * C * xz.element_size()

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
        tf = t(lambda: ops.causal_conv1d_channel_last(x, w, b, True))
    y = ops.causal_conv1d_channel_last(x, w, b, True)

    def bwd():
        x.grad = w.grad = b.grad = None
        y.backward(gy, retain_graph=True)
    tb = t(bwd)
    print(f"{os.path.basename(os.environ.get('BDLRU_LIB', 'libbdlru.so')):32s} B={B} T={T} C={C} {str(dt)[6:]}: "
          f"fwd {tf:.3f} ms {2 * E / tf / 1e6:.0f} GB/s | bwd {tb:.3f} ms {3 * E / tb / 1e6:.0f} GB/s", flush=True)


if __name__ == "__main__":
    main()
