#!/bin/bash
# N-GPU box: configs[3], NCCL scatter vs peer-memory pull (pipelined)
mkdir -p gpurun_out
N=${1:-2}
i=0
IFS=","; for cfg in ${CFGS:-nccl 1,p2p 1,p2p 2}; do unset IFS
set -- $cfg; sc=$1; sb=$2; i=$((i+1))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$i bench.py --gpus $N --workload c4 --steps 10 --warmup 3 --scatter $sc --subbatches $sb > gpurun_out/c4_c4_${N}gpu_${sc}_sb$sb.json 2> gpurun_out/c4_c4_${N}gpu_${sc}_sb$sb.err; echo "$sc sb=$sb rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/c4_c4_${N}gpu_${sc}_sb$sb.json').read().strip().splitlines()[-1])
    c=d['configs3']
    print('$sc sb=$sb', 'pieces', c['sub_batches'], 'value', round(d['value']), 'ms', round(c['ms_per_step'],3), 'scatter', round(c['scatter_ms'],3), 'solve', round(c['solve_ms'],3), 'gather', round(c['gather_ms'],3), 'identical', c['gathered_equals_single_gpu_solve'])
except Exception as e:
    print('$sc sb=$sb FAILED', e); import subprocess; print(open('gpurun_out/c4_c4_${N}gpu_${sc}_sb$sb.err').read()[-1500:])
PY
done
