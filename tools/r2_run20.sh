#!/bin/bash
# ncu capture of the tensor-core pruning screen (last pass of a 100 k pruning call), with source correlation
mkdir -p gpurun_out
python tools/run_c4.py 100000 > gpurun_out/c4_plain.log 2>&1 && tail -1 gpurun_out/c4_plain.log &&
ncu --set full --clock-control none --import-source on -k regex:gram_tc_kernel -s 10 -c 1 -f -o gpurun_out/r2b_gram python tools/run_c4.py 100000 > gpurun_out/ncu_c4.log 2>&1
tail -2 gpurun_out/ncu_c4.log
ls -la gpurun_out/*.ncu-rep
