#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_clash_gpu.py -x -q -k "cell or compact or level or bitmask or adversarial or mask_matches" 2>&1 | tail -8
python tools/prof_cell.py 10000000 5
python tools/prof_cell.py 10000000 5 q7
FC_CLASH_LEVELS=32,64 python tools/prof_cell.py 10000000 5
FC_CLASH_LEVELS=16 python tools/prof_cell.py 10000000 5
FC_CLASH_LEVELS=24 python tools/prof_cell.py 10000000 5
FC_CLASH_LEVELS=149 python tools/prof_cell.py 10000000 5
python tools/prof_cell.py 2000000 4 > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:clash_cell_kernel -s 6 -c 2 -f -o gpurun_out/r2_cell_v3 python tools/prof_cell.py 2000000 4 > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/ncu1.log
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -s 24 -c 8 --csv --log-file gpurun_out/r2_cell_v3_launches.csv python tools/prof_cell.py 2000000 4 > gpurun_out/ncu2.log 2>&1
grep -v "^==" gpurun_out/r2_cell_v3_launches.csv | awk -F'","' '{print $5, $13, $15}' | tail -20
