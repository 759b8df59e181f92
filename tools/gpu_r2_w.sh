#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ba.py -m gpu -q -x 2>&1 | tail -3
GMMLOC_B200_LIB=$PWD/gmmloc_b200/lib/libgmmloc_b200_baprof.so python tools/prof_ba.py 8 2>&1 | grep "BA profile" | tail -1 | sed 's/|/\n/g'
timeout 900 python tools/bench_ba.py 2> gpurun_out/r2w_bench_ba.err | tee gpurun_out/r2w_bench_ba.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in ('keyframes', 'points', 'windows', 'cluster', 'gpu_ms_per_call', 'cpu_ms_per_window_1core', 'dpose_max')})"
