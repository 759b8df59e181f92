#!/bin/bash
mkdir -p gpurun_out
run() {  # name, lib
  GMMLOC_B200_LIB=$2 timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2s_bench_$1.json 2> gpurun_out/r2s_bench_$1.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2s_bench_$1.json').read().strip().splitlines()[-1])
    print('$1', 'c3 value', round(d['value']), {k: round(v, 3) for k, v in d['kernel_ms'].items()}, d['roofline']['kernel'])
except Exception as e:
    print('$1', 'FAILED', e)
PY
}
run default ""
run t256 $PWD/gmmloc_b200/lib/libgmmloc_b200_t256.so
