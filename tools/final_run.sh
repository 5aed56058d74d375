python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r01g.json 2> gpurun_out/bench_r01g.err; echo bench rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; echo ref rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01g.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras --e2e-poses 1000000 > gpurun_out/ncu_launch_g.log 2>&1
echo ncu launches rc=$?
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
