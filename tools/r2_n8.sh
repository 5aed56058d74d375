#!/bin/bash
# round 2: the 8-GPU lines (one box): bench.py at N = 8 (C3 weak / strong scaling, end to end, C4 sharded) and the C4 tool
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/c4_sharded.py 200000 > gpurun_out/c4_sharded_n$N.json 2> gpurun_out/c4_sharded_n$N.err; echo "c4 rc=$?"
cat gpurun_out/c4_sharded_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_r02b_n$N.json 2> gpurun_out/bench_r02b_n$N.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_r02b_n$N.json
