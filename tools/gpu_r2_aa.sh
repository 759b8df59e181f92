#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "associate or grid or tensor or frames_solve_c2 or c3_full or c5_full" 2>&1 | tail -3
bash tools/gpu_r2_z.sh
