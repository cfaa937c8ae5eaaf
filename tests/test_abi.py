"""CPU-side checks of the drop-in boundary: libbdlru.so builds for sm_100a without a GPU, loads, and exports every
symbol include/bdlru.h declares (no compute calls here)."""
import os
import re

from tests.conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "bdlru.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(bdlru_[a-z0-9_]+)\s*\(", src))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from datamining_recblr_b200 import _lib, build
    path = build.build()
    assert os.path.exists(path)
    lib = _lib.load()
    assert _lib.MISSING == []
    declared = _declared()
    assert declared, "no declarations parsed from include/bdlru.h"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib.cdll, name), name
    assert lib.bdlru_version() == _lib.ABI_VERSION
    assert isinstance(lib.bdlru_last_error(), bytes)


def test_graft_entry_build():
    import __graft_entry__ as g
    g.build()


def test_bad_arguments_return_error_codes_without_a_gpu():
    """Argument validation happens before any CUDA call, so it is testable here: bad sizes -> BDLRU_ERR_INVALID and a
    message from bdlru_last_error()."""
    from datamining_recblr_b200 import _lib
    lib = _lib.load()
    rc = lib.bdlru_scan_fwd(None, None, None, 0, 0, None)
    assert rc == 1
    assert lib.bdlru_last_error()


def test_ctypes_signatures_match_header_parameter_counts():
    """Every ctypes `argtypes` list in _lib.SIGNATURES has as many entries as the C declaration in include/bdlru.h has
    parameters (a bdlru_view passed by value counts as one) — guards against the binding drifting from the header."""
    from datamining_recblr_b200 import _lib
    src = open(os.path.join(ROOT, "include", "bdlru.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = re.findall(r"\b(bdlru_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S)
    assert len(decls) == len(_lib.SIGNATURES)
    for name, params in decls:
        params = params.strip()
        n = 0 if params in ("", "void") else len([p for p in params.split(",") if p.strip()])
        assert n == len(_lib.SIGNATURES[name][1]), (name, n, len(_lib.SIGNATURES[name][1]))
