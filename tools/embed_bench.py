"""Times the front end LayerNorm(dropout(table[ids])) forward at the training shape (random ids over a large bf16 table):
algorithmic bytes per token 8 + 2 * D * esize.  BDLRU_LIB selects a variant library.

    python tools/embed_bench.py [tokens] [n_items] [D]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from datamining_recblr_b200 import ops  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192 * 200
    items = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
    D = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    dev = "cuda"
    table = (torch.randn(items, D, device=dev) * 0.02).to(torch.bfloat16)
    ids = torch.randint(1, items, (n,), device=dev)
    g, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)

    def t(fn, k=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        evs = []
        for _ in range(k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            evs.append((s, e))
        torch.cuda.synchronize()
        ts = sorted(s.elapsed_time(e) for s, e in evs)
        return ts[len(ts) // 2]

    with torch.no_grad():
        tf = t(lambda: ops.embed_layernorm(ids, table, g, b, 1e-12, 0.2, 7))
    by = n * (8 + 2 * D * 2)
    print(f"{os.path.basename(os.environ.get('BDLRU_LIB', 'libbdlru.so')):32s} tokens={n} items={items} D={D} bf16: "
          f"fwd {tf:.3f} ms {by / tf / 1e6:.0f} GB/s", flush=True)


if __name__ == "__main__":
    main()
