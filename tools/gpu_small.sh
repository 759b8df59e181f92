#!/bin/bash
# stage times of small device-resident batches (the chunks of the host pipeline)
mkdir -p gpurun_out
for n in 148 296 592 888 1184 4096; do
  timeout 300 python bench.py --frames $n --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/small_$n.json 2> gpurun_out/small_$n.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/small_$n.json').read().strip().splitlines()[-1])
    print($n, 'ms_per_step', round(d['ms_per_step'],4), {k: round(v, 4) for k, v in d['kernel_ms'].items()}, 'e2e', round(d['e2e']['value']))
except Exception as e:
    print($n, 'FAILED', e)
PY
done
