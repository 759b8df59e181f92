#!/bin/bash
mkdir -p gpurun_out
run() {  # name, lib, gn
  GMMLOC_B200_LIB=$2 timeout 600 python bench.py --steps 10 --warmup 3 --gn $3 --no-cpu-baseline --no-extras > gpurun_out/r2f_bench_$1.json 2> gpurun_out/r2f_bench_$1.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2f_bench_$1.json').read().strip().splitlines()[-1])
    print('$1', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), {k: round(v, 3) for k, v in d['kernel_ms'].items()})
except Exception as e:
    print('$1', 'FAILED', e)
PY
}
run mixed "" mixed
run world256 "" world
for v in gmmloc_b200/lib/libgmmloc_b200_*.so; do
  n=$(basename $v .so); n=${n#libgmmloc_b200_}
  run world_$n $PWD/$v world
done
timeout 900 python -m pytest tests/test_gpu_ba.py -m gpu -q -x 2>&1 | tail -3
timeout 900 python tools/bench_ba.py 2> gpurun_out/r2f_bench_ba.err | tee gpurun_out/r2f_bench_ba.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in ('keyframes', 'points', 'windows', 'cluster', 'gpu_ms_per_call', 'cpu_ms_per_window_1core', 'iters')})"
