#!/bin/bash
# 2 GPUs: weak-scaling bench with overlapped bitmask all-gather, strong-scaling leg, e2e with NUMA-placed pinned buffers
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -q 0x10de $d/vendor 2>/dev/null; then echo "$d $(cat $d/numa_node)"; fi; done >> gpurun_out/r2_topo.txt
python -c "import os; print('affinity', sorted(os.sched_getaffinity(0)))" >> gpurun_out/r2_topo.txt
cat /sys/devices/system/node/node*/cpulist >> gpurun_out/r2_topo.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus ${NG:-2} --steps 20 --warmup 3 > gpurun_out/r2_bench_n${NG:-2}.json 2> gpurun_out/r2_bench_n${NG:-2}.err
echo "rc=$?"; tail -5 gpurun_out/r2_bench_n${NG:-2}.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n${NG:-2}.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print('strong',d['strong_scaling']); print('e2e',d['e2e']['value'],d['e2e']['h2d_gbs_pinned_measured'],d['e2e']['ms_per_step_each']); print(d['device_resident_pose7'])
PY
tail -12 gpurun_out/r2_topo.txt
