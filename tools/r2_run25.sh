#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_prune_gpu.py -x -q -k "flavours or c4_20k or conventions or all_similar or sharded" > gpurun_out/pytest25.log 2>&1
grep -v "^  File" gpurun_out/pytest25.log | tail -6
FC_PRUNE_TRACE=1 python tools/run_c4.py 200000 > gpurun_out/c4_trace.log 2>&1
grep -E "fc_prune: total|kept=|pass k=(2|1) " gpurun_out/c4_trace.log | tail -4
python tools/run_c4.py 200000 2>&1 | tail -1
