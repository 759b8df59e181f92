#!/bin/bash
mkdir -p gpurun_out
L=$PWD/gmmloc_b200/lib
run() {  # name workload lib env...
  local name=$1 w=$2 lib=$3; shift 3
  env "$@" GMMLOC_B200_LIB=$lib timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/lv_${name}_$w.json 2> gpurun_out/lv_${name}_$w.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/lv_${name}_$w.json').read().strip().splitlines()[-1])
    print('$name', '$w', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), {k: round(v, 4) for k, v in d['kernel_ms'].items()}, d.get('frames_numeric_failure'))
except Exception as e:
    print('$name', '$w', 'FAILED', e)
PY
}
for w in ${WL:-c2 c3 c5}; do
  run t25_h05 $w $L/libgmmloc_b200_dbg.so GMMLOC_GRID_TMAX0=25 GMMLOC_GRID_HSCALE=0.5
  run t30_h05 $w $L/libgmmloc_b200_dbg.so GMMLOC_GRID_TMAX0=30 GMMLOC_GRID_HSCALE=0.5
  run t30_h06 $w $L/libgmmloc_b200_dbg.so GMMLOC_GRID_TMAX0=30 GMMLOC_GRID_HSCALE=0.6
  run t25_h04 $w $L/libgmmloc_b200_dbg.so GMMLOC_GRID_TMAX0=25 GMMLOC_GRID_HSCALE=0.4
  run t25_100_h05 $w $L/libgmmloc_b200_dbg.so GMMLOC_GRID_TMAX0=25 GMMLOC_GRID_TMAX1=100 GMMLOC_GRID_HSCALE=0.5
done
python tools/time_map_create.py 2>&1 | tail -5
GMMLOC_B200_LIB=$L/libgmmloc_b200_dbg.so GMMLOC_GRID_TMAX0=25 GMMLOC_GRID_HSCALE=0.5 python tools/time_map_create.py 2>&1 | tail -5
