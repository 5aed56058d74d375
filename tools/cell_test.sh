python -m pytest tests/test_clash_gpu.py -x -q 2>&1 | tail -2
FC_CLASH_MODE=1 python -m pytest tests/test_cyclical3_embed_gpu.py tests/test_string_embed_gpu.py -x -q 2>&1 | tail -2
python bench.py --steps 10 --no-cpu --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['e2e']['value'])"
