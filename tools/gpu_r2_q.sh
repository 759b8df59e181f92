#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pipelined or deterministic or two_contexts" 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2q_bench.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e'].get('matches_device_arm'), {k: round(v, 3) for k, v in d['kernel_ms'].items()}, d.get('pcie'))
PY
timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2q_bench_c3.json 2> gpurun_out/r2q_bench_c3.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2q_bench_c3.json').read().strip().splitlines()[-1])
print('c3 value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e'].get('matches_device_arm'))
PY
