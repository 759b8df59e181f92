set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/s6_pytest.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s6_bench_ref.json 2> gpurun_out/s6_bench_ref.err
timeout 400 python bench.py > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err
timeout 300 python bench.py --workload c5 --no-cpu-baseline --no-extras > gpurun_out/s6_bench_c5.json 2> gpurun_out/s6_bench_c5.err
timeout 300 python bench.py --workload c3 --no-cpu-baseline --no-extras > gpurun_out/s6_bench_c3.json 2> gpurun_out/s6_bench_c3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s6_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/s6_launch_run.log 2>&1
tail -3 gpurun_out/s6_pytest.log; for f in gpurun_out/s6_bench.json gpurun_out/s6_bench_c5.json gpurun_out/s6_bench_c3.json; do cut -c1-400 $f; echo; done
