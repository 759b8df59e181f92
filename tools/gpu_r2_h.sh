#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stream or world_frame_kernel_matches or agrees_with_the_other" > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -15 gpurun_out/r2h_pytest.log
run() {  # name, gn
  timeout 300 python bench.py --steps 10 --warmup 3 --gn $2 --no-cpu-baseline --no-extras > gpurun_out/r2h_bench_$1.json 2> gpurun_out/r2h_bench_$1.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2h_bench_$1.json').read().strip().splitlines()[-1])
    print('$1', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), {k: round(v, 3) for k, v in d['kernel_ms'].items()}, 'numfail', d.get('frames_numeric_failure'))
except Exception as e:
    print('$1', 'FAILED', e)
PY
}
run mixed mixed
run stream stream
