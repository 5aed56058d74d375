#!/bin/bash
# round 2, first GPU call: clash tests, short bench, launch list + full capture of the cell-list screen
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests/test_clash_gpu.py -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
echo "bench rc=$?"; tail -5 gpurun_out/r2_bench1.err; cat gpurun_out/r2_bench1.json | cut -c1-3000
