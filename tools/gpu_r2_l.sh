#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stream or world_frame_kernel_matches or agrees_with_the_other" 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --gn stream --no-cpu-baseline --no-extras > gpurun_out/r2l_bench_stream.json 2> gpurun_out/r2l_bench_stream.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2l_bench_stream.json').read().strip().splitlines()[-1])
print('stream value', round(d['value']), 'e2e', round(d['e2e']['value']), {k: round(v, 3) for k, v in d['kernel_ms'].items()}, 'numfail', d.get('frames_numeric_failure'))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 30 -c 40 --csv --log-file gpurun_out/r2l_launches.csv python bench.py --steps 3 --warmup 3 --gn stream --no-cpu-baseline --no-extras > gpurun_out/r2l_launch_run.log 2>&1
grep -E "frame_" gpurun_out/r2l_launches.csv | awk -F'","' '{print $5, $NF}' | sed 's/gmmloc::<unnamed>:://; s/(.*)//' | sort | uniq -c | head -30
