#!/bin/bash
set -x
mkdir -p gpurun_out
python tools/prof_cell.py 2000000 4 > gpurun_out/prof_plain.log 2>&1 && cat gpurun_out/prof_plain.log &&
ncu --set full --clock-control none --import-source on -k regex:clash_cell_kernel -s 9 -c 3 -f -o gpurun_out/r2_cell python tools/prof_cell.py 2000000 4 > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/ncu1.log
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -s 27 -c 9 --csv --log-file gpurun_out/r2_cell_launches.csv python tools/prof_cell.py 2000000 4 > gpurun_out/ncu2.log 2>&1
cat gpurun_out/r2_cell_launches.csv | tail -30
FC_CLASH_LEVELS=149 python tools/prof_cell.py 10000000 5
FC_CLASH_LEVELS=32 python tools/prof_cell.py 10000000 5
FC_CLASH_LEVELS=16,48 python tools/prof_cell.py 10000000 5
FC_CLASH_LEVELS=16,32,64 python tools/prof_cell.py 10000000 5
FC_CLASH_LEVELS=32,64,96 python tools/prof_cell.py 10000000 5
python tools/prof_cell.py 10000000 5
python tools/prof_cell.py 10000000 5 q7
