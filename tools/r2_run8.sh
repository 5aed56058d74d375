#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cyclical3_embed_gpu.py tests/test_dist_gloo.py -x -q 2>&1 | tail -5
FC_CLASH_TRACE=1 timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err
echo "rc=$?"; grep -v "^fc_clash_batch" gpurun_out/r2_bench_full.err | tail -12
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_full.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print(d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['issue'])
print('e2e',d['e2e']['value'], d['e2e']['xf64_api']['value']); print(d['cpu_baseline'])
for k,v in d['other_workloads'].items(): print(k, {a:b for a,b in v.items() if a not in ('note','roofline','conventions')})
PY
