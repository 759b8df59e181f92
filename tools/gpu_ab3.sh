#!/bin/bash
# last A/B of the round: GPU tests of the frame-GN paths on the default build, then default / variant / default bench lines
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -x > gpurun_out/ab3_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/ab3_pytest.log
L=$PWD/gmmloc_b200/lib
for v in "" $L/libgmmloc_b200_*.so ""; do
  name=default; [ -n "$v" ] && { name=$(basename $v .so); name=${name#libgmmloc_b200_}; }
  GMMLOC_B200_LIB=$v timeout 100 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/ab3_$name.json 2> gpurun_out/ab3_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab3_$name.json').read().strip().splitlines()[-1])
    print('$name', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), {k: round(v, 4) for k, v in d['kernel_ms'].items()}, d.get('frames_numeric_failure'))
except Exception as e:
    print('$name', 'FAILED', e)
PY
done
