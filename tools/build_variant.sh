#!/bin/bash
# Builds a compile-time variant of the library next to the default one (A/B timing on the GPU box):
#   tools/build_variant.sh t192 -DGMMLOC_GNW_THREADS2=192     ->  gmmloc_b200/lib/libgmmloc_b200_t192.so
# Use it with  GMMLOC_B200_LIB=gmmloc_b200/lib/libgmmloc_b200_t192.so python bench.py ...
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
cp -r "$root/gmmloc_b200/build" "$tmp/build" 2>/dev/null || mkdir -p "$tmp/build"
GMMLOC_NVCC_EXTRA="$*" python - "$root" "$tmp" "$name" <<'PY'
import sys, os, subprocess, shutil
from pathlib import Path
root, tmp, name = Path(sys.argv[1]), Path(sys.argv[2]), sys.argv[3]
sys.path.insert(0, str(root))
from gmmloc_b200 import build as b
b.OBJ_DIR = tmp / "obj"
b.LIB = root / "gmmloc_b200" / "lib" / f"libgmmloc_b200_{name}.so"
print(b.build(force=True))
PY
rm -rf "$tmp"
