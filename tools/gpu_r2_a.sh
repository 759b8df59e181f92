#!/bin/bash
# round-2 GPU call A: full -m gpu suite, smoke, default bench (1 GPU)
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2a_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?" >> gpurun_out/r2a_bench.err
tail -5 gpurun_out/r2a_pytest.log; tail -2 gpurun_out/r2a_smoke.log; tail -3 gpurun_out/r2a_bench.err; head -c 1500 gpurun_out/r2a_bench.json
