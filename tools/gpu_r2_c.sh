#!/bin/bash
# round-2 GPU call C: full -m gpu suite (incl. K3w), A/B timing of the frame-GN kernels, ncu of K3w
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -15 gpurun_out/r2c_pytest.log
for gn in mixed world; do
  timeout 600 python bench.py --steps 10 --warmup 3 --gn $gn --no-cpu-baseline --no-extras > gpurun_out/r2c_bench_$gn.json 2> gpurun_out/r2c_bench_$gn.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2c_bench_$gn.json').read().strip().splitlines()[-1])
print('$gn', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['kernel_ms'])
PY
done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:frame_gn_world --launch-skip 3 -c 1 -f -o gpurun_out/r2c_gnw python bench.py --steps 1 --warmup 3 --gn world --no-cpu-baseline --no-extras > gpurun_out/r2c_ncu.log 2>&1
ls -la gpurun_out/r2c_gnw.ncu-rep
