#!/bin/bash
# 2-GPU box: multirank parity test, configs[3] (scatter / solve / gather) and the weak-scaling c2 line
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x > gpurun_out/mgpu_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/mgpu_pytest.log
tail -5 gpurun_out/mgpu_pytest.log
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload c4 --steps 10 --warmup 3 > gpurun_out/mgpu_bench_c4_${N}gpu.json 2> gpurun_out/mgpu_bench_c4_${N}gpu.err; echo "c4 rc=$?"
cut -c1-1800 gpurun_out/mgpu_bench_c4_${N}gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-extras > gpurun_out/mgpu_bench_c2_${N}gpu.json 2> gpurun_out/mgpu_bench_c2_${N}gpu.err; echo "c2 rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/mgpu_bench_c2_${N}gpu.json').read().strip().splitlines()[-1])
print('c2 x$N value', round(d['value']), 'e2e', round(d['e2e']['value']), d['kernel_ms'], d.get('pcie'))
PY
