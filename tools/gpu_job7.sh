#!/bin/bash
# round-2 GPU job 7 (1 GPU): full GPU suite, default bench, evidence captures
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/j7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j7_pytest.log
tail -8 gpurun_out/j7_pytest.log
timeout 300 python tools/train_bench.py > gpurun_out/j7_train_bench.log 2>&1; tail -2 gpurun_out/j7_train_bench.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/j7_bench.json 2> gpurun_out/j7_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/j7_bench_ref.json 2> gpurun_out/j7_bench_ref.err; echo "ref rc=$?"
# launch list of the default bench (shares) and one full capture of an MLP launch INSIDE the sustained loop (7th image)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/j7_bench_launches.csv \
  python bench.py --only --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/j7_bench_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 13 -c 1 -o gpurun_out/j7_mlp_sustained \
  python bench.py --only --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/j7_mlp_ncu.log 2>&1
timeout 300 python tools/parity_report.py > gpurun_out/j7_parity_report.json 2> gpurun_out/j7_parity.err
tail -c 300 gpurun_out/j7_bench.json
