#!/bin/bash
# round-2 GPU job 8 (1 GPU): training forward emits the backward operands (no recompute); mesh path on grow-only buffers
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/j8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j8_pytest.log
tail -12 gpurun_out/j8_pytest.log
timeout 300 python tools/train_bench.py > gpurun_out/j8_train_bench.log 2>&1; tail -2 gpurun_out/j8_train_bench.log
NM_TRAIN_DIRECT_GB=0 timeout 300 python tools/train_bench.py > gpurun_out/j8_train_bench_sub.log 2>&1; tail -1 gpurun_out/j8_train_bench_sub.log
timeout 300 python tools/mc_bench.py --lego > gpurun_out/j8_mc_lego.log 2>&1; tail -5 gpurun_out/j8_mc_lego.log
timeout 600 python bench.py --workload mesh --only --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/j8_mesh.json 2> gpurun_out/j8_mesh.err; echo "mesh rc=$?"
tail -c 1500 gpurun_out/j8_mesh.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/j8_train_launches.csv python tools/train_profile.py > gpurun_out/j8_train_ncu.log 2>&1
