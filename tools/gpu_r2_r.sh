#!/bin/bash
mkdir -p gpurun_out
export GMMLOC_B200_LIB=$PWD/gmmloc_b200/lib/libgmmloc_b200_dbg.so
GMMLOC_E2E_TRACE=1 PROBE_CHUNKS=0 timeout 300 python tools/probe_e2e.py 2>&1 | tail -40
GMMLOC_E2E_TRACE=1 GMMLOC_E2E_CHUNK=888 timeout 300 python tools/probe_e2e.py 888 2>&1 | tail -8
