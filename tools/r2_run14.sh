#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tfd_keepfirst_gpu.py tests/test_string_embed_gpu.py tests/test_reference_fixtures_gpu.py tests/test_prune_gpu.py --durations=8 -x -q > gpurun_out/pytest14.log 2>&1
grep -v "^  File" gpurun_out/pytest14.log | tail -16
FC_PRUNE_TRACE=1 python tools/run_c4.py 200000 > gpurun_out/c4_trace.log 2>&1
grep -E "fc_prune: total|kept=|upload_rows" gpurun_out/c4_trace.log
python tools/run_c4.py 200000 2>&1 | tail -1
python tools/run_c1.py > gpurun_out/c1_plain.log 2>&1; tail -9 gpurun_out/c1_plain.log
