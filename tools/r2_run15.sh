#!/bin/bash
mkdir -p gpurun_out
python tools/run_c1.py > gpurun_out/c1_plain.log 2>&1; tail -9 gpurun_out/c1_plain.log
timeout 900 python -m pytest tests/test_tfd_keepfirst_gpu.py tests/test_string_embed_gpu.py tests/test_reference_fixtures_gpu.py -x -q > gpurun_out/pytest15.log 2>&1
grep -v "^  File" gpurun_out/pytest15.log | tail -8
