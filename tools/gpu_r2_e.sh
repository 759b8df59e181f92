#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ba.py tests/test_gpu_parity.py -m gpu -q -k "ba or c5_full or fp64_kernel_and" > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -6 gpurun_out/r2e_pytest.log
timeout 900 python tools/bench_ba.py > gpurun_out/r2e_bench_ba.jsonl 2> gpurun_out/r2e_bench_ba.err; cat gpurun_out/r2e_bench_ba.jsonl | cut -c1-330
timeout 900 ncu --set full --import-source on --clock-control none -k regex:frame_gn_world --launch-skip 3 -c 1 -f -o gpurun_out/r2e_gnw256 python bench.py --steps 1 --warmup 3 --gn world --no-cpu-baseline --no-extras > gpurun_out/r2e_ncu256.log 2>&1
GMMLOC_B200_LIB=$PWD/gmmloc_b200/lib/libgmmloc_b200_t192.so timeout 900 ncu --set full --import-source on --clock-control none -k regex:frame_gn_world --launch-skip 3 -c 1 -f -o gpurun_out/r2e_gnw192 python bench.py --steps 1 --warmup 3 --gn world --no-cpu-baseline --no-extras > gpurun_out/r2e_ncu192.log 2>&1
ls -la gpurun_out/*.ncu-rep
