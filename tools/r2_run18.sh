#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_prune_gpu.py -x -q -k "sharded or matches_oracle or conventions" > gpurun_out/pytest18.log 2>&1
grep -v "^  File" gpurun_out/pytest18.log | tail -12
FC_PRUNE_TRACE=1 python tools/run_c4.py 200000 > gpurun_out/c4_trace.log 2>&1
grep -E "fc_prune: total|kept=|upload_rows" gpurun_out/c4_trace.log | tail -3
python tools/run_c4.py 200000 2>&1 | tail -1
