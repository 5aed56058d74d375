FC_CLASH_MODE=0 python -m pytest tests/test_clash_gpu.py -x -q 2>&1 | tail -2
FC_CLASH_MODE=0 python bench.py --steps 5 --no-cpu --no-extras --e2e-poses 1000000 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
