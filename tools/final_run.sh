python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r01d.json 2> gpurun_out/bench_r01d.err; echo bench rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras --e2e-poses 1000000 > gpurun_out/ncu_launch_d.log 2>&1
echo ncu rc=$?
