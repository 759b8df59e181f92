#!/bin/bash
# GPU tests + short bench lines of every single-GPU config with the default build
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/chk_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/chk_pytest.log
for spec in "c2 4096" "c2 296" "c2 888" "c3 0" "c5 0"; do
  set -- $spec; w=$1; n=$2
  fr=""; [ "$n" != 0 ] && fr="--frames $n"
  timeout 300 python bench.py --workload $w $fr --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/chk_${w}_$n.json 2> gpurun_out/chk_${w}_$n.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/chk_${w}_$n.json').read().strip().splitlines()[-1])
    print('$w', '$n', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), {k: round(v, 4) for k, v in d['kernel_ms'].items()}, d.get('frames_numeric_failure'), 'assoc frac', round(d['roofline_assoc']['frac'], 3))
except Exception as e:
    print('$w', '$n', 'FAILED', e)
PY
done
