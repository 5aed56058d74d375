python -m pytest tests/test_clash_gpu.py -x -q 2>&1 | tail -5
FC_CLASH_MODE=1 python -m pytest tests/test_string_embed_gpu.py tests/test_cyclical_embed_gpu.py tests/test_cyclical3_embed_gpu.py -x -q 2>&1 | tail -3
for m in 0 1; do echo -n "MODE=$m: "; FC_CLASH_MODE=$m python bench.py --steps 5 --warmup 3 --no-cpu --no-extras --e2e-poses 10000000 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['e2e']['value'], d['config']['pass_fraction'], d['config']['fp64_rechecks'])"; done
