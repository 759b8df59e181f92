#!/bin/bash
# round-2 evidence run: launch list of the default bench command + ncu --set full captures of the step's kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prof_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/prof_launch_run.log 2>&1
cap() {  # name, kernel regex, launch-skip
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:$2 --launch-skip $3 -c 1 -f -o gpurun_out/prof_$1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/prof_ncu_$1.log 2>&1
}
cap gns frame_gn_stream 3
cap pack frame_pack 3
cap scan_tiles excl_scan_tiles 3
cap grid assoc_grid_kernel 3
cap scatter grid_scatter 3
cap hist grid_hist 3
ls -la gpurun_out/prof_*.ncu-rep
