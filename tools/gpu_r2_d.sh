#!/bin/bash
# round-2 GPU call D: full -m gpu suite, A/B timing of the frame-GN kernels (incl. compile-time variants)
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -8 gpurun_out/r2d_pytest.log
run() {  # name, lib, gn
  GMMLOC_B200_LIB=$2 timeout 600 python bench.py --steps 10 --warmup 3 --gn $3 --no-cpu-baseline --no-extras > gpurun_out/r2d_bench_$1.json 2> gpurun_out/r2d_bench_$1.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2d_bench_$1.json').read().strip().splitlines()[-1])
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
if [ -n "$NCU_VARIANT" ]; then
  GMMLOC_B200_LIB=$NCU_VARIANT timeout 900 ncu --set full --import-source on --clock-control none -k regex:frame_gn_world --launch-skip 3 -c 1 -f -o gpurun_out/r2d_gnw python bench.py --steps 1 --warmup 3 --gn world --no-cpu-baseline --no-extras > gpurun_out/r2d_ncu.log 2>&1
fi
