#!/bin/bash
mkdir -p gpurun_out
timeout 100 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize_late_r2.py > gpurun_out/racecheck_late.txt 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/racecheck_late.txt
timeout 95 compute-sanitizer --tool memcheck python tools/sanitize_late_r2.py > gpurun_out/memcheck_late.txt 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck_late.txt
