#!/bin/bash
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/memcheck_smoke.py > gpurun_out/r2v_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2v_memcheck.log
tail -6 gpurun_out/r2v_memcheck.log
MEMCHECK_LIGHT=1 timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/memcheck_smoke.py > gpurun_out/r2v_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2v_racecheck.log
grep -c "Race reported\|hazard" gpurun_out/r2v_racecheck.log; tail -6 gpurun_out/r2v_racecheck.log
