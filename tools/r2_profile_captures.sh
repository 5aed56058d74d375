#!/bin/bash
# round 2 final profiler captures: tensor-core pruning screen (FP16 build), string-embed sweep kernels
mkdir -p gpurun_out
python tools/run_c4.py 100000 > gpurun_out/c4_plain.log 2>&1 && tail -1 gpurun_out/c4_plain.log &&
ncu --set full --clock-control none --import-source on -k regex:gram_tc_kernel -s 10 -c 1 -f -o gpurun_out/r2c_gram python tools/run_c4.py 100000 > gpurun_out/ncu_c4.log 2>&1
tail -1 gpurun_out/ncu_c4.log
python tools/run_c1.py > gpurun_out/c1_plain.log 2>&1 && tail -2 gpurun_out/c1_plain.log &&
ncu --set full --clock-control none --import-source on -k regex:"tfd_block_matrix_kernel|tfd_block_resolve_kernel" -s 4 -c 2 -f -o gpurun_out/r2c_tfd python tools/run_c1.py > gpurun_out/ncu_c1.log 2>&1
tail -1 gpurun_out/ncu_c1.log
ls -la gpurun_out/*.ncu-rep
