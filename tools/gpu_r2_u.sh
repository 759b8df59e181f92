#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:local_ba --launch-skip 1 -c 1 -f -o gpurun_out/r2u_ba8 python tools/prof_ba.py 8 > gpurun_out/r2u_ncu_ba.log 2>&1
ls -la gpurun_out/r2u_ba8.ncu-rep
