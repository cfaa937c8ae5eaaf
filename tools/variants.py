"""Builds tuning variants of libbdlru.so that differ from the in-tree build in the -D flags of ONE source file, into
datamining_recblr_b200/variants/ (git-ignored, travels to the GPU box).  Select one at run time with BDLRU_LIB=<path>.

    python tools/variants.py add_ln.cu r1:-DBDLRU_ADD_LN_ROWS=1 r2:-DBDLRU_ADD_LN_ROWS=2,-DBDLRU_ADD_LN_BPSM=8
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from datamining_recblr_b200 import build as B  # noqa: E402

OUT = os.path.join(ROOT, "datamining_recblr_b200", "variants")


def main():
    src, specs = sys.argv[1], sys.argv[2:]
    B.build()
    os.makedirs(OUT, exist_ok=True)
    for spec in specs:
        name, _, flags = spec.partition(":")
        obj = os.path.join(OUT, f"{src[:-3]}_{name}.o")
        subprocess.run([B.NVCC, *B.FLAGS, "-DBDLRU_TUNING", *[f for f in flags.split(",") if f], "-c",
                        os.path.join(B.CSRC, src), "-o", obj], check=True)
        objs = [obj if s == src else os.path.join(B.OBJ, s[:-3] + ".o") for s in B._sources()]
        lib = os.path.join(OUT, f"libbdlru_{src[:-3]}_{name}.so")
        subprocess.run([B.NVCC, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"], check=True)
        print(lib, flush=True)


if __name__ == "__main__":
    main()
