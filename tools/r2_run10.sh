#!/bin/bash
# round 2 housekeeping captures: tensor-core pruning screen (6-issuer build), torsion scan, launch lists of the small workloads
set -x
mkdir -p gpurun_out
python tools/run_c4.py 100000 > gpurun_out/c4_plain.log 2>&1 && tail -2 gpurun_out/c4_plain.log &&
ncu --set full --clock-control none --import-source on -k regex:gram_tc_kernel -s 10 -c 1 -f -o gpurun_out/r2_gram python tools/run_c4.py 100000 > gpurun_out/ncu_c4.log 2>&1
tail -2 gpurun_out/ncu_c4.log
python tools/run_c5.py > gpurun_out/c5_plain.log 2>&1 && tail -1 gpurun_out/c5_plain.log &&
ncu --set full --clock-control none --import-source on -k regex:torsion_scan_kernel -s 2 -c 1 -f -o gpurun_out/r2_torsion python tools/run_c5.py > gpurun_out/ncu_c5.log 2>&1
tail -2 gpurun_out/ncu_c5.log
python tools/run_c1.py > gpurun_out/c1_plain.log 2>&1 && tail -1 gpurun_out/c1_plain.log &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/r2_c1_launches.csv python tools/run_c1.py > /dev/null 2>&1
python -m pytest tests/test_csearch_gpu.py tests/test_tfd_prune_gpu.py tests/test_cyclical_embed_gpu.py -x -q 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_small_launches.csv python -m pytest tests/test_csearch_gpu.py tests/test_tfd_prune_gpu.py tests/test_cyclical_embed_gpu.py -x -q > /dev/null 2>&1
echo done
