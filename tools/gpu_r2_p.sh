#!/bin/bash
# round-2 evidence run: launch list of the default bench command + ncu --set full captures of the step's kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2p_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2p_launch_run.log 2>&1
cap() {  # name, kernel regex, launch-skip
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:$2 --launch-skip $3 -c 1 -f -o gpurun_out/r2p_$1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2p_ncu_$1.log 2>&1
}
cap gns frame_gn_stream 3
cap pack frame_pack 3
cap classify frame_classify 3
cap grid assoc_grid_kernel 3
cap scatter grid_scatter 3
cap hist grid_hist 3
ls -la gpurun_out/r2p_*.ncu-rep
for v in t160 t128; do
GMMLOC_B200_LIB=$PWD/gmmloc_b200/lib/libgmmloc_b200_$v.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2p_bench_$v.json 2> gpurun_out/r2p_bench_$v.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2p_bench_$v.json').read().strip().splitlines()[-1])
    print('$v', 'value', round(d['value']), {k: round(x, 3) for k, x in d['kernel_ms'].items()})
except Exception as e:
    print('$v', 'FAILED', e)
PY
done
