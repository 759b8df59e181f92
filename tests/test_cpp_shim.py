"""The header-only C++ shim (include/gmmloc_b200.hpp) compiles against the C ABI (CPU) and, on the GPU
box, a C++ caller written against the reference's names gets the known answers."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "cpp" / "shim_test.cpp"
LIBDIR = ROOT / "gmmloc_b200" / "lib"


def _build(out):
    from gmmloc_b200 import capi
    capi.lib()
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", str(ROOT / "include"), str(SRC), "-o", str(out), "-L", str(LIBDIR),
           "-lgmmloc_b200", f"-Wl,-rpath,{LIBDIR}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)


def test_shim_compiles_and_links(tmp_path):
    _build(tmp_path / "shim_test")


@pytest.mark.gpu
def test_shim_known_answers(tmp_path):
    exe = tmp_path / "shim_test"
    _build(exe)
    r = subprocess.run([str(exe), str(tmp_path / "lattice.gmm")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
