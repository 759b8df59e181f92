set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/s6_pytest.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s6_bench_ref.json 2> gpurun_out/s6_bench_ref.err
timeout 400 python bench.py > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:frame_gn_mixed -s 2 -c 1 -o gpurun_out/s6_gnm -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/s6_ncu.log 2>&1
tail -3 gpurun_out/s6_pytest.log; cat gpurun_out/s6_bench.json | cut -c1-600
