python -m pytest tests/test_clash_gpu.py -x -q 2>&1 | tail -3
for m in 1; do echo -n "MODE=$m: "; FC_CLASH_MODE=$m python bench.py --steps 5 --warmup 3 --no-cpu --no-extras --e2e-poses 10000000 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['e2e']['value'], d['config']['pass_fraction'], d['config']['fp64_rechecks'])"; done
FC_CLASH_MODE=1 ncu --set full --clock-control none --import-source on -k regex:clash_cell -c 1 -o gpurun_out/prof_cell_v2 -f python bench.py --steps 1 --warmup 3 --no-cpu --no-extras --poses 2000000 --e2e-poses 100000 > gpurun_out/ncu_cell.log 2>&1
echo ncu rc=$?
