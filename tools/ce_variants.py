"""Builds tuning variants of libbdlru.so that differ only in the fraction of softmax exponentials the fused-CE kernels
compute on the FMA pipe (common.cuh ex2_mixed), into datamining_recblr_b200/variants/, and (with --run, on a GPU box)
runs the CE parity tests and tools/ce_bench.py against each through BDLRU_LIB.

    python tools/ce_variants.py            # build here (nvcc cross-compiles)
    python tools/ce_variants.py --run      # on the GPU box: test + time each variant
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from datamining_recblr_b200 import build as B  # noqa: E402

OUT = os.path.join(ROOT, "datamining_recblr_b200", "variants")
# (dQ mask, dE mask, forward mask): bit i set = column i of every 32-column chunk uses the polynomial
M4, M6, M8, M11, M12, M14 = 0x80808080, 0x84108410, 0x88888888, 0x92492492, 0x92929292, 0x95295295
MASKS = [(0, 0, 0), (M8, M6, M4), (M11, M8, M6), (M12, M11, M8), (M14, M12, M12)]
VARIANTS = {"q%d_e%d_f%d" % tuple(bin(m).count("1") for m in t): t for t in MASKS}


def build_all():
    os.makedirs(OUT, exist_ok=True)
    for name, (mq, me, mf) in VARIANTS.items():
        objdir = os.path.join(OUT, "obj_" + name)
        os.makedirs(objdir, exist_ok=True)
        defs = ["-DBDLRU_TUNING", f"-DBDLRU_CE_DQ_POLY_MASK={mq:#x}u", f"-DBDLRU_CE_DE_POLY_MASK={me:#x}u", f"-DBDLRU_CE_FWD_POLY_MASK={mf:#x}u"]
        objs = []
        for src in B._sources():
            obj = os.path.join(objdir, src[:-3] + ".o")
            special = src in ("fullsort.cu", "fullsort_bwd.cu")
            base_obj = os.path.join(OUT, "obj_q0_e0_f0", src[:-3] + ".o")
            if not special and name != "q0_e0_f0":
                objs.append(base_obj)
                continue
            subprocess.run([B.NVCC, *B.FLAGS, *(defs if special else []), "-c", os.path.join(B.CSRC, src), "-o", obj], check=True)
            objs.append(obj)
        lib = os.path.join(OUT, f"libbdlru_{name}.so")
        subprocess.run([B.NVCC, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"], check=True)
        print(name, [bin(m).count("1") for m in (mq, me, mf)], lib, flush=True)


def run_all():
    names = list(VARIANTS)
    for name in names:
        env = dict(os.environ, BDLRU_LIB=os.path.join(OUT, f"libbdlru_{name}.so"))
        print(f"===== {name}", flush=True)
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ce_bench.py"), "8192", "1000000"], env=env, timeout=200)
    for name in sys.argv[sys.argv.index("--run") + 1:] or names[-1:]:   # parity tests: named variants, default the most aggressive
        env = dict(os.environ, BDLRU_LIB=os.path.join(OUT, f"libbdlru_{name}.so"))
        r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_fullsort.py"), "-q", "-x", "-m", "gpu",
                            "-k", "ce", "-p", "no:cacheprovider"], env=env, timeout=600, capture_output=True, text=True)
        print(f"pytest {name}:", r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
        if r.returncode:
            print(r.stdout[-3000:])


if __name__ == "__main__":
    run_all() if "--run" in sys.argv else build_all()
