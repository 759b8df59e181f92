#!/bin/bash
# Association-grid options on the BASELINE workloads (debug-env build: tools/build_variant.sh dbg -DGMMLOC_DEBUG_ENV):
#   CONFIGS="name:ENV=val,ENV=val ..." bash tools/gpu_levels.sh        (WL="c2 c3 c5" selects the workloads)
# The late round-2 sweeps: T0 in 9 / 16 / 20 / 25 / 30 / 36, T1 in 36 / 64 / 100 / 144, cell scale 0.4 .. 0.9.
mkdir -p gpurun_out
L=$PWD/gmmloc_b200/lib
CONFIGS=${CONFIGS:-"default: t36_h075:GMMLOC_GRID_TMAX0=36,GMMLOC_GRID_HSCALE=0.75 t25_h075:GMMLOC_GRID_HSCALE=0.75 t30:GMMLOC_GRID_TMAX0=30 t25_100:GMMLOC_GRID_TMAX1=100 t25_h05:GMMLOC_GRID_HSCALE=0.5"}
for w in ${WL:-c2 c3 c5}; do
  for cfg in $CONFIGS; do
    name=${cfg%%:*}; envs=${cfg#*:}
    env ${envs//,/ } GMMLOC_B200_LIB=$L/libgmmloc_b200_dbg.so timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/lv_${name}_$w.json 2> gpurun_out/lv_${name}_$w.err
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/lv_${name}_$w.json').read().strip().splitlines()[-1])
    print('$name', '$w', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), {k: round(v, 4) for k, v in d['kernel_ms'].items()}, d.get('frames_numeric_failure'))
except Exception as e:
    print('$name', '$w', 'FAILED', e)
PY
  done
done
