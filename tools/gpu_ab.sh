#!/bin/bash
# variants x batch sizes (device-resident stage times + e2e), after the GPU tests on the default build
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != 1 ]; then
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/ab_pytest.log
fi
L=$PWD/gmmloc_b200/lib
for n in ${SIZES:-4096 296 888 148}; do
for v in "" $L/libgmmloc_b200_*.so; do
  name=default; [ -n "$v" ] && { name=$(basename $v .so); name=${name#libgmmloc_b200_}; }
  GMMLOC_B200_LIB=$v timeout 300 python bench.py --frames $n --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/ab2_${name}_$n.json 2> gpurun_out/ab2_${name}_$n.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab2_${name}_$n.json').read().strip().splitlines()[-1])
    print('$name', $n, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), {k: round(v, 4) for k, v in d['kernel_ms'].items()}, d.get('frames_numeric_failure'))
except Exception as e:
    print('$name', $n, 'FAILED', e)
PY
done
done
