python -m pytest tests -m gpu -x -q 2>&1 | tail -3
FC_CLASH_MODE=1 ncu --set full --clock-control none --import-source on -k regex:clash_cell -c 1 -o gpurun_out/prof_cell_final -f python bench.py --steps 1 --warmup 3 --no-cpu --no-extras --poses 2000000 --e2e-poses 100000 > gpurun_out/ncu_cell_final.log 2>&1
echo ncu cell rc=$?
python bench.py > gpurun_out/bench_r01e.json 2> gpurun_out/bench_r01e.err; echo bench rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; echo ref rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01e.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras --e2e-poses 1000000 > gpurun_out/ncu_launch_e.log 2>&1
echo ncu launches rc=$?
