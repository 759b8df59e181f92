#!/bin/bash
# one pass over every variant library in gmmloc_b200/lib (default first and last), after the GPU tests on the default build
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != 1 ]; then
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/ab_pytest.log
fi
run() {  # name, lib
  local t0=$(date +%s)
  GMMLOC_B200_LIB=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/ab_bench_$1.json 2> gpurun_out/ab_bench_$1.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab_bench_$1.json').read().strip().splitlines()[-1])
    print('$1', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), {k: round(v, 4) for k, v in d['kernel_ms'].items()}, d.get('frames_numeric_failure'), 'wall', $(date +%s)-$t0)
except Exception as e:
    print('$1', 'FAILED', e)
PY
}
L=$PWD/gmmloc_b200/lib
run warm ""
run default ""
for v in $L/libgmmloc_b200_*.so; do
  n=$(basename $v .so); n=${n#libgmmloc_b200_}
  run $n $v
done
run default2 ""
