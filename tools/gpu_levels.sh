#!/bin/bash
# cooperative level loop of the sorted grid scan: GPU tests, then C2 / C3 / C5 with and without it, then level thresholds (debug-env build)
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != 1 ]; then
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/lv_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/lv_pytest.log
fi
L=$PWD/gmmloc_b200/lib
run() {  # name workload lib env...
  local name=$1 w=$2 lib=$3; shift 3
  env "$@" GMMLOC_B200_LIB=$lib timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/lv_${name}_$w.json 2> gpurun_out/lv_${name}_$w.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/lv_${name}_$w.json').read().strip().splitlines()[-1])
    print('$name', '$w', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), {k: round(v, 4) for k, v in d['kernel_ms'].items()}, 'flagged', round(d.get('rerank_flagged_frac', -1), 4), d.get('frames_numeric_failure'))
except Exception as e:
    print('$name', '$w', 'FAILED', e)
PY
}
for w in ${WL:-c2 c3 c5}; do
  run default $w "" X=1
  run nocoop $w $L/libgmmloc_b200_nocoop.so X=1
  run t25_144 $w $L/libgmmloc_b200_dbg.so GMMLOC_GRID_TMAX0=25
  run t16_144 $w $L/libgmmloc_b200_dbg.so GMMLOC_GRID_TMAX0=16
  run t16_64 $w $L/libgmmloc_b200_dbg.so GMMLOC_GRID_TMAX0=16 GMMLOC_GRID_TMAX1=64
  run t9_36 $w $L/libgmmloc_b200_dbg.so GMMLOC_GRID_TMAX0=9 GMMLOC_GRID_TMAX1=36
done
