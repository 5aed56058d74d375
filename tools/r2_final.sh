#!/bin/bash
# round 2 end-of-session run on one B200: the GPU test suite, smoke(), both bench arms, the launch list and one full ncu
# capture of the headline kernel (each profiler pass only after its command has run clean without ncu)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"
grep -v "^  File" gpurun_out/pytest_final.log | tail -18
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -8 gpurun_out/smoke_final.log
python bench.py > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02b_ref.json 2>/dev/null; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras --e2e-poses 1000000 > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
python tools/run_c1.py > gpurun_out/c1_final.log 2>&1; tail -12 gpurun_out/c1_final.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/r02b_c1_launches.csv python tools/run_c1.py > /dev/null 2>&1
echo "c1 launches rc=$?"
