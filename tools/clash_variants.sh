python -m pytest tests/test_clash_gpu.py -x -q 2>&1 | tail -2
for u in 11 13 15; do
  echo -n "UNR=$u: "; FC_CLASH_UNROLL=$u python bench.py --steps 5 --warmup 3 --no-cpu --no-extras --e2e-poses 1000000 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['executed_frac'])"
done
