python -m pytest tests/test_clash_gpu.py tests/test_dist_gloo.py -x -q 2>&1 | tail -2
FC_CLASH_TRACE=1 python bench.py --steps 3 --no-cpu --no-extras 2>gpurun_out/trace.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['h2d_bound_poses_per_s'])"
grep fc_clash_batch gpurun_out/trace.err | tail -2
