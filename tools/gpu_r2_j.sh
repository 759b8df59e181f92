#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 30 -c 40 --csv --log-file gpurun_out/r2j_launches.csv python bench.py --steps 3 --warmup 3 --gn stream --no-cpu-baseline --no-extras > gpurun_out/r2j_launch_run.log 2>&1
grep -E "frame_|assoc_|grid_" gpurun_out/r2j_launches.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | head -30
timeout 900 ncu --set full --import-source on --clock-control none -k regex:frame_gn_stream --launch-skip 3 -c 1 -f -o gpurun_out/r2j_gns python bench.py --steps 1 --warmup 3 --gn stream --no-cpu-baseline --no-extras > gpurun_out/r2j_ncu.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:frame_pack --launch-skip 3 -c 1 -f -o gpurun_out/r2j_pack python bench.py --steps 1 --warmup 3 --gn stream --no-cpu-baseline --no-extras > gpurun_out/r2j_ncu_pack.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
