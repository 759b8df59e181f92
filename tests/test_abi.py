"""The C-ABI library loads and exports every symbol include/gmmloc_b200.h declares; host-side logic
that needs no device; and the product refuses (loudly) to compute without a GPU — there is no CPU
fallback and nothing in the package touches oracle/."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    txt = (ROOT / "include" / "gmmloc_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"GMMLOC_API\s+[\w\s\*]+?\b(gmmloc_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from gmmloc_b200 import capi
    names = _declared()
    assert len(names) >= 25
    L = capi.lib()
    for n in names:
        assert hasattr(L, n), n
    assert sorted(capi.SYMBOLS) == names  # the binding covers the header, nothing more, nothing less


def test_version_and_status_strings():
    from gmmloc_b200 import capi
    L = capi.lib()
    assert b"sm_100a" in L.gmmloc_version()
    assert L.gmmloc_status_string(capi.ECUDA) == b"CUDA error / no device"


def test_default_params_are_the_shipped_values(oracle):
    import gmmloc_b200 as g
    p, o = g.default_params(), oracle.default_params()
    for name, _ in g.Params._fields_:
        assert getattr(p, name) == getattr(o, name), name
    assert p.tri_lambda2 == 400.0 and p.gn_iters == 10 and p.mdist2_gate == 9.0
    assert p.tri_str_thresh == float(np.float32(0.0064))


def test_sigma2_inv_matches_reference_float_recurrence(oracle):
    import gmmloc_b200 as g
    for o in range(-1, 10):
        assert g.sigma2_inv(o) == oracle.sigma2_inv(o)


def test_struct_layouts_match_the_header():
    from gmmloc_b200 import capi
    assert ctypes.sizeof(capi.Camera) == 48
    assert ctypes.sizeof(capi.Params) == 72
    assert ctypes.sizeof(capi.Stats) == 32


def test_argument_validation_needs_no_device():
    import gmmloc_b200 as g
    from gmmloc_b200 import capi
    L = capi.lib()
    h = ctypes.c_void_p()
    assert L.gmmloc_map_create(None, None, 0, 1e-4, 0, ctypes.byref(h)) == capi.EINVAL
    assert L.gmmloc_map_size(None) == 0
    assert L.gmmloc_ctx_synchronize(None) == capi.EINVAL
    assert b"null" in L.gmmloc_last_error()
    with pytest.raises(g.GmmlocError) as ei:
        g.GMM.load("/nonexistent/file.gmm")
    assert ei.value.status == capi.EIO


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import gmmloc_b200 as g
    from gmmloc_b200 import capi
    with pytest.raises(g.GmmlocError) as ei:
        g.Context(0)
    assert ei.value.status == capi.ECUDA
    with pytest.raises(g.GmmlocError) as ei:
        g.GMM(np.zeros((4, 3)), np.tile(np.eye(3).ravel(), (4, 1)))
    assert ei.value.status == capi.ECUDA


def test_product_never_imports_the_oracle():
    for p in (ROOT / "gmmloc_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h", ".hpp", ".cpp"):
            txt = p.read_text()
            assert "pyoracle" not in txt and "gmmloc_oracle" not in txt, p
    for p in (ROOT / "include").glob("*"):
        assert "oracle" not in p.read_text().lower(), p
