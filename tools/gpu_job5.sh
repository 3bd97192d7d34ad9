#!/bin/bash
# round-2 GPU job 5 (1 GPU): fused forward-recompute in the training backward; MC stage timing; inference regression check
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_train.py tests/test_gpu_wide_parity.py -m gpu -q -x > gpurun_out/j5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j5_pytest.log
tail -12 gpurun_out/j5_pytest.log
python tools/train_bench.py > gpurun_out/j5_train_bench.log 2>&1; tail -8 gpurun_out/j5_train_bench.log
NM_TRAIN_LAYERWISE=1 python tools/train_bench.py > gpurun_out/j5_train_bench_layerwise.log 2>&1; tail -4 gpurun_out/j5_train_bench_layerwise.log
python tools/mc_bench.py > gpurun_out/j5_mc_bench.log 2>&1; tail -6 gpurun_out/j5_mc_bench.log
python bench.py --only --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/j5_lego.json 2> gpurun_out/j5_lego.err; tail -c 700 gpurun_out/j5_lego.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/j5_train_launches.csv python tools/train_profile.py > gpurun_out/j5_train_ncu.log 2>&1
