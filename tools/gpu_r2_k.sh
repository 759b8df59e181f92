#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/diag_status.py 4096 > gpurun_out/r2k_diag.log 2>&1; tail -20 gpurun_out/r2k_diag.log
timeout 300 python bench.py --steps 10 --warmup 3 --gn stream --no-cpu-baseline --no-extras > gpurun_out/r2k_bench_stream.json 2> gpurun_out/r2k_bench_stream.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2k_bench_stream.json').read().strip().splitlines()[-1])
print('stream value', round(d['value']), 'e2e', round(d['e2e']['value']), {k: round(v, 3) for k, v in d['kernel_ms'].items()}, 'numfail', d.get('frames_numeric_failure'))
PY
