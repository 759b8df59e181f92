#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/full_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/full_pytest.log
tail -5 gpurun_out/full_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/full_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/full_bench.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), d['kernel_ms'], d.get('parity'), d['roofline']['kernel'], d['roofline']['frac'])
PY
for w in c3 c5; do
timeout 600 python bench.py --workload $w --no-extras > gpurun_out/full_bench_$w.json 2> gpurun_out/full_bench_$w.err
python - <<PY
import json
d=json.loads(open('gpurun_out/full_bench_$w.json').read().strip().splitlines()[-1])
print('$w', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['kernel_ms'], d.get('parity'), d['roofline']['kernel'], 'numfail', d.get('frames_numeric_failure'))
PY
done
