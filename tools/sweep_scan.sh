#!/bin/bash
# GPU-box helper: rebuild with different scan shapes and print the per-kernel times.
cd "$(dirname "$0")/.."
while read -r cfg; do
  [ -z "$cfg" ] && continue
  GMMLOC_NVCC_EXTRA="$cfg" python -m gmmloc_b200.build --force > /tmp/build.log 2>&1 || { echo "build failed: $cfg"; tail -3 /tmp/build.log; continue; }
  echo "== $cfg"
  python bench.py --steps 5 --warmup 3 --frames 2048 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['kernel_ms'], d['e2e']['matches_device_arm'])"
done <<'CFG'
-DGMMLOC_SCAN_P=8 -DGMMLOC_SCAN_THREADS=256 -DGMMLOC_SCAN_MINB=2
-DGMMLOC_SCAN_P=8 -DGMMLOC_SCAN_THREADS=128 -DGMMLOC_SCAN_MINB=4
-DGMMLOC_SCAN_P=8 -DGMMLOC_SCAN_THREADS=256 -DGMMLOC_SCAN_MINB=2 -DGMMLOC_CHUNK=32
-DGMMLOC_SCAN_P=12 -DGMMLOC_SCAN_THREADS=128 -DGMMLOC_SCAN_MINB=3
-DGMMLOC_SCAN_P=16 -DGMMLOC_SCAN_THREADS=128 -DGMMLOC_SCAN_MINB=2
-DGMMLOC_SCAN_P=6 -DGMMLOC_SCAN_THREADS=256 -DGMMLOC_SCAN_MINB=2
-DGMMLOC_SCAN_P=4 -DGMMLOC_SCAN_THREADS=256 -DGMMLOC_SCAN_MINB=3
-DGMMLOC_SCAN_P=10 -DGMMLOC_SCAN_THREADS=192 -DGMMLOC_SCAN_MINB=2
CFG
python -m gmmloc_b200.build --force > /dev/null 2>&1
