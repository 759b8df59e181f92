#!/bin/bash
# GPU-box helper: rebuild frame_gn.cu with different launch bounds and print the per-kernel times.
cd "$(dirname "$0")/.."
for cfg in "128 3" "128 4" "256 2" "256 1" "64 6"; do
  set -- $cfg
  GMMLOC_NVCC_EXTRA="-DGMMLOC_GN_THREADS=$1 -DGMMLOC_GN_MINB=$2" python -m gmmloc_b200.build --force > /dev/null 2>&1 || { echo "build failed $cfg"; continue; }
  echo "== GN threads=$1 minb=$2"
  python bench.py --steps 5 --warmup 3 --frames 2048 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['kernel_ms'], d['frames_numeric_failure'])"
done
python -m gmmloc_b200.build --force > /dev/null 2>&1
