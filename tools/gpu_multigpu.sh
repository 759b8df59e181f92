#!/bin/bash
# N-GPU box: multirank parity test (N = 2), configs[3] (scatter / solve / gather; NCCL and peer-memory pull) and the weak-scaling c2 line
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L | head -8
if [ "$N" = 2 ]; then
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x > gpurun_out/mgpu_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/mgpu_pytest.log
tail -5 gpurun_out/mgpu_pytest.log
fi
i=0
for cfg in ${CFGS:-nccl:1 p2p:1 p2p:2}; do
sc=${cfg%%:*}; sb=${cfg##*:}; i=$((i+1))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$i bench.py --gpus $N --workload c4 --steps 10 --warmup 3 --scatter $sc --subbatches $sb > gpurun_out/mgpu_c4_${N}gpu_${sc}_sb$sb.json 2> gpurun_out/mgpu_c4_${N}gpu_${sc}_sb$sb.err; echo "c4 $sc sb=$sb rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/mgpu_c4_${N}gpu_${sc}_sb$sb.json').read().strip().splitlines()[-1])
    c=d['configs3']
    print('c4 x$N $sc pieces', c['sub_batches'], 'value', round(d['value']), 'ms', round(c['ms_per_step'],3), 'scatter', round(c['scatter_ms'],3), 'solve', round(c['solve_ms'],3), 'gather', round(c['gather_ms'],3), 'identical', c['gathered_equals_single_gpu_solve'])
except Exception as e:
    print('c4 $sc sb=$sb FAILED', e); print(open('gpurun_out/mgpu_c4_${N}gpu_${sc}_sb$sb.err').read()[-1500:])
PY
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 --no-extras > gpurun_out/mgpu_bench_c2_${N}gpu.json 2> gpurun_out/mgpu_bench_c2_${N}gpu.err; echo "c2 rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/mgpu_bench_c2_${N}gpu.json').read().strip().splitlines()[-1])
print('c2 x$N value', round(d['value']), 'e2e', round(d['e2e']['value']), d['kernel_ms'], d.get('pcie'))
PY
