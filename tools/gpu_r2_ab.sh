#!/bin/bash
# N-GPU box: configs[3] with and without the pipelined scatter
mkdir -p gpurun_out
N=${1:-2}
for sb in 1 0 2 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$sb bench.py --gpus $N --workload c4 --steps 10 --warmup 3 --subbatches $sb > gpurun_out/r2ab_c4_${N}gpu_sb$sb.json 2> gpurun_out/r2ab_c4_${N}gpu_sb$sb.err; echo "sb=$sb rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2ab_c4_${N}gpu_sb$sb.json').read().strip().splitlines()[-1])
    c=d['configs3']
    print('sb=$sb', 'pieces', c['sub_batches'], 'value', round(d['value']), 'ms', round(c['ms_per_step'],3), 'scatter', round(c['scatter_ms'],3), 'solve', round(c['solve_ms'],3), 'gather', round(c['gather_ms'],3), 'identical', c['gathered_equals_single_gpu_solve'])
except Exception as e:
    print('sb=$sb FAILED', e)
PY
done
