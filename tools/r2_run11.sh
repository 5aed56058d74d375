#!/bin/bash
# round 2, session 2: C4 stage trace (before the host-side rework) and C1 stage timing
mkdir -p gpurun_out
FC_PRUNE_TRACE=1 python tools/run_c4.py 200000 > gpurun_out/c4_trace.log 2>&1
tail -40 gpurun_out/c4_trace.log
python tools/run_c1.py > gpurun_out/c1_plain.log 2>&1; tail -5 gpurun_out/c1_plain.log
nproc; lscpu | grep -E "Model name|Socket|Core|Thread" 
