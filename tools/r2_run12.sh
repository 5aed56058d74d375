#!/bin/bash
# round 2, session 2: host-side rework of the pruning driver + bit-matrix TFD sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tfd_keepfirst_gpu.py tests/test_string_embed_gpu.py tests/test_reference_fixtures_gpu.py tests/test_prune_gpu.py tests/test_refining_gpu.py -x -q > gpurun_out/pytest12.log 2>&1
grep -v "^  File" gpurun_out/pytest12.log | tail -12
FC_PRUNE_TRACE=1 python tools/run_c4.py 200000 > gpurun_out/c4_trace.log 2>&1
grep -E "fc_prune: total|kept=" gpurun_out/c4_trace.log
python tools/run_c4.py 200000 2>&1 | tail -2
python tools/run_c1.py > gpurun_out/c1_plain.log 2>&1; tail -14 gpurun_out/c1_plain.log
