#!/bin/bash
# timeline of the host pipeline of gmmloc_frames_solve (debug-env build): last call's per-chunk trace
mkdir -p gpurun_out
GMMLOC_B200_LIB=$PWD/gmmloc_b200/lib/libgmmloc_b200_dbg.so GMMLOC_E2E_TRACE=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/trace_bench.json 2> gpurun_out/trace_bench.err
grep "^chunk" gpurun_out/trace_bench.err | tail -14
