#!/bin/bash
# round-2 GPU job 24 (1 GPU): validation of the final build — full suite, smoke, default bench + reference arm, evidence captures
set -x
mkdir -p gpurun_out
python -m nerfmeshes_b200.build > gpurun_out/j24_build.log 2>&1 || { tail -5 gpurun_out/j24_build.log; exit 9; }
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/j24_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j24_pytest.log
tail -4 gpurun_out/j24_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/j24_smoke.log 2>&1; tail -2 gpurun_out/j24_smoke.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/j24_bench.json 2> gpurun_out/j24_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/j24_bench_ref.json 2> gpurun_out/j24_bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/j24_bench_launches.csv \
  python bench.py --only --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/j24_bench_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/j24_train_launches.csv python tools/train_profile.py 4096 > gpurun_out/j24_train_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 13 -c 1 -o gpurun_out/j24_mlp_sustained \
  python bench.py --only --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/j24_mlp_ncu.log 2>&1
timeout 300 python tools/power_trace.py > gpurun_out/j24_power_trace.json 2> gpurun_out/j24_power.err
timeout 300 python tools/train_bench.py > gpurun_out/j24_train_bench.log 2>&1; tail -1 gpurun_out/j24_train_bench.log
tail -c 300 gpurun_out/j24_bench.json
