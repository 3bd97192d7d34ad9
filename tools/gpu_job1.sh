#!/bin/bash
# round-2 GPU job 1 (1 GPU): parity tests, default bench, launch lists + one full capture of the AABB sampler
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/j1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j1_pytest.log
python bench.py --steps 4 --warmup 3 > gpurun_out/j1_bench.json 2> gpurun_out/j1_bench.err; echo "bench rc=$?" >> gpurun_out/j1_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/j1_buff_launches.csv \
  python bench.py --workload buff --only --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/j1_buff_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:aabb_kernel -s 1 -c 1 -o gpurun_out/j1_aabb \
  python bench.py --workload buff --only --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/j1_aabb_ncu.log 2>&1
tail -5 gpurun_out/j1_pytest.log; tail -c 1500 gpurun_out/j1_bench.json; tail -3 gpurun_out/j1_bench.err
