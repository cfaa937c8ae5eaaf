"""TEST INFRASTRUCTURE ONLY — recipe for `oracle/_ref/` (git-ignored, travels to the GPU box with the snapshot).

The reference's BD-LRU path is pure Python + Triton (RecBLR.py, parallel_scan.py): nothing to compile, but its Triton
kernel — "the kernel to beat" of SURVEY Appendix C — can only launch on a GPU, and /root/reference does not exist on the
GPU box.  This recipe packs the two UNMODIFIED reference files, byte for byte from where they lie under /root/reference,
into `oracle/_ref/reference_py.tar.gz` (never into the repository's history), with their sha256 recorded beside it.
`load_gpu_reference()` unpacks the blob into a temporary directory at run time and imports it with the RecBole stub —
used ONLY by `tests/test_gpu_vs_reference.py` (parity against the real Triton scan on the B200) and by bench.py's
`vs_triton` leg (speed of the reference layer on the same box).  The product never imports it.

    python -m oracle.build_ref      # in the build container; __graft_entry__.build() calls build() too
"""
import hashlib
import importlib
import io
import json
import os
import sys
import tarfile
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("RECBLR_REFERENCE_ROOT", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
BLOB = os.path.join(OUT_DIR, "reference_py.tar.gz")
FILES = ("RecBLR.py", "parallel_scan.py")
_STUB = os.path.join(HERE, "recbole_stub")


def build():
    """Packs the reference files when /root/reference is present; a no-op (keeping any existing blob) elsewhere."""
    if not all(os.path.isfile(os.path.join(REF_ROOT, f)) for f in FILES):
        return os.path.exists(BLOB)
    os.makedirs(OUT_DIR, exist_ok=True)
    digests = {}
    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode="w:gz") as tar:
        for f in FILES:
            path = os.path.join(REF_ROOT, f)
            digests[f] = hashlib.sha256(open(path, "rb").read()).hexdigest()
            info = tar.gettarinfo(path, arcname=f)
            info.mtime = 0
            with open(path, "rb") as fh:
                tar.addfile(info, fh)
    with open(BLOB, "wb") as fh:
        fh.write(buf.getvalue())
    with open(os.path.join(OUT_DIR, "MANIFEST.json"), "w") as fh:
        json.dump({"source": REF_ROOT, "sha256": digests}, fh, indent=1)
    return True


def available():
    return os.path.isfile(BLOB)


_loaded = None


def load_gpu_reference():
    """(RecBLR module, parallel_scan module) of the unmodified reference, Triton scan intact (needs a CUDA device to
    launch).  `causal_conv1d` is absent in this image, so RecBLR.py:8-11 takes its own F.conv1d fallback (line 185)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("oracle/_ref/reference_py.tar.gz is missing: run `python -m oracle.build_ref` where "
                           "/root/reference exists")
    tmp = tempfile.mkdtemp(prefix="recblr_ref_")
    with tarfile.open(BLOB) as tar:
        tar.extractall(tmp, filter="data")
    man = json.load(open(os.path.join(OUT_DIR, "MANIFEST.json")))
    for f, want in man["sha256"].items():
        got = hashlib.sha256(open(os.path.join(tmp, f), "rb").read()).hexdigest()
        assert got == want, f"{f}: unpacked reference differs from the recorded sha256"
    for p in (_STUB, tmp):
        if p not in sys.path:
            sys.path.insert(0, p)
    for name in ("RecBLR", "parallel_scan"):
        sys.modules.pop(name, None)
    ps = importlib.import_module("parallel_scan")
    mod = importlib.import_module("RecBLR")
    _loaded = (mod, ps)
    return _loaded


if __name__ == "__main__":
    print("packed" if build() else "reference tree not found; nothing packed", BLOB)
