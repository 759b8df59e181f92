#!/bin/bash
# GPU-box helper: rebuild frame_gn_mixed.cu with different compile-time knobs (SWEEP_CFGS: space-separated lists of
# comma-separated -D definitions) and print the per-kernel times of the default bench workload.
cd "$(dirname "$0")/.."
CFGS=${SWEEP_CFGS:-GMMLOC_GNM_THREADS2=192,GMMLOC_GNM_MAXREG2=168}
for cfg in $CFGS; do
  defs=""
  for d in ${cfg//,/ }; do defs="$defs -D$d"; done
  touch gmmloc_b200/csrc/frame_gn_mixed.cu
  GMMLOC_NVCC_EXTRA="$defs" python -m gmmloc_b200.build > /dev/null 2>&1 || { echo "build failed $cfg"; continue; }
  echo "== $cfg"
  for rep in 1 2; do
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['kernel_ms'], d['frames_numeric_failure'])"
  done
done
touch gmmloc_b200/csrc/frame_gn_mixed.cu
python -m gmmloc_b200.build > /dev/null 2>&1
