import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_maps():
    out = {}
    for name in ("v1", "v2"):
        z = np.load(GOLDEN / f"{name}_gmm.npz")
        out[name] = dict(mean=z["mean"], cov=z["cov"], flags=z["flags"], sha256=str(z["sha256"]), nbytes=int(z["nbytes"]))
    return out


@pytest.fixture(scope="session")
def gt_v1_03():
    return np.load(GOLDEN / "v1_03_gt_sync.npz")["gt"]


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def gpu():
    """The product package with a live CUDA context (only -m gpu tests request this)."""
    import gmmloc_b200 as g
    ctx = g.Context(0)
    yield g, ctx
    ctx.close()
