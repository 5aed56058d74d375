#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_prune_gpu.py -x -q > gpurun_out/pytest26.log 2>&1; echo "pytest rc=$?"
grep -v "^  File" gpurun_out/pytest26.log | tail -6
FC_PRUNE_TRACE=1 python tools/run_c4.py 200000 > gpurun_out/c4_trace.log 2>&1
grep -E "fc_prune: total|kept=|pass k=" gpurun_out/c4_trace.log | tail -15
python tools/run_c4.py 200000 2>&1 | tail -1
FC_PRUNE_CULL=0 python tools/run_c4.py 200000 2>&1 | tail -1
