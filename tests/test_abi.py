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


def test_host_pipeline_chunk_schedule_needs_no_device():
    """gmmloc_frames_solve_chunks: the chunk schedule of the host pipeline is pure host arithmetic (api.cu,
    frames_solve_chunk_starts).  Every schedule covers the batch exactly; large batches run the graded schedule
    (2, 3, 4, ... frames per SM, growth capped at twice the chunk option), small ones one chunk."""
    import gmmloc_b200 as g
    assert g.frames_solve_chunks(4096) == [296, 444, 592, 740, 888, 1136]  # the C2 step measured in DESIGN.md 5
    assert g.frames_solve_chunks(1) == [1] and g.frames_solve_chunks(1023) == [1023]
    assert g.frames_solve_chunks(1100) == [1100]           # 1100 - 888 < half a chunk: merged
    assert g.frames_solve_chunks(2149) == [888, 1261]      # below 24 frames per SM: uniform chunks, short rest merged
    for sms in (1, 4, 132, 148, 160):
        for B in (1, 7, 1023, 1024, 1500, 3551, 3552, 3700, 4096, 8192, 50000):
            for chunk in (0, 100, 888, 3000):
                for taper in (0, 1):
                    for short_first in (0, 1):
                        sz = g.frames_solve_chunks(B, sms, chunk, taper, short_first)
                        assert sum(sz) == B and all(x > 0 for x in sz), (sms, B, chunk, taper, short_first, sz)
    big = g.frames_solve_chunks(50000)
    assert big[0] == 296 and big[1] == 444 and max(big[:-1]) == 1776  # capped at 12 frames per SM
    assert all(b >= a for a, b in zip(big[:-2], big[1:-1]))
    assert g.frames_solve_chunks(50000, chunk=400)[5] == 800  # the option caps the growth at twice its value
    assert g.frames_solve_chunks(4096, taper=0) == [888, 888, 888, 888, 544]
    from gmmloc_b200 import capi
    assert capi.lib().gmmloc_frames_solve_chunks(0, 148, 0, 1, 0, None, 0) == -1
    assert capi.lib().gmmloc_frames_solve_chunks(4096, 148, 0, 1, 0, None, 0) == 6  # count only
