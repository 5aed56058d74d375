#!/bin/bash
# final check of the last changes (three-stage screen, want_structures) and the N = 1 bench line of the final build
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_prune_gpu.py tests/test_refining_gpu.py tests/test_rot_corr_gpu.py -x -q > gpurun_out/pytest_final2.log 2>&1; echo "pytest rc=$?"
grep -v "^  File" gpurun_out/pytest_final2.log | tail -5
python bench.py > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err; echo "bench rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
