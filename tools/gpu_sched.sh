#!/bin/bash
# sweep of chunk schedules of the host pipeline (debug-env build, GMMLOC_E2E_SCHEDULE)
mkdir -p gpurun_out
i=0
while read -r sch; do
  i=$((i+1))
  GMMLOC_B200_LIB=$PWD/gmmloc_b200/lib/libgmmloc_b200_dbg.so GMMLOC_E2E_SCHEDULE=$sch timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/sched_$i.json 2> gpurun_out/sched_$i.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/sched_$i.json').read().strip().splitlines()[-1])
    print('${sch:-default}', 'e2e', round(d['e2e']['value']), 'value', round(d['value']), d['e2e'].get('matches_device_arm'))
except Exception as e:
    print('${sch:-default}', 'FAILED', e)
PY
done <<'S'

296,444,592,740,888,1136
296,592,592,888,888,840
296,592,888,1036,1284
296,592,888,1184,1136
296,444,592,888,888,988
148,296,444,592,740,888,988
296,592,888,888,888,544
296,592,1184,2024
592,888,1184,1432
296,592,544,888,888,888
444,592,740,888,1432
S
