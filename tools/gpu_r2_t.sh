#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "associate or grid or tensor or frames_solve_c2 or c3_full" 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2t_bench.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), {k: round(v, 3) for k, v in d['kernel_ms'].items()}, d['brute_force_scan']['assignments_identical_to_default'])
PY
for w in c3 c5; do
timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2t_bench_$w.json 2> gpurun_out/r2t_bench_$w.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2t_bench_$w.json').read().strip().splitlines()[-1])
print('$w value', round(d['value']), {k: round(v, 3) for k, v in d['kernel_ms'].items()}, d['brute_force_scan']['assignments_identical_to_default'])
PY
done
