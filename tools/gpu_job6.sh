#!/bin/bash
# round-2 GPU job 6 (1 GPU): fused data-gradient chain (training backward), mesh path stage timing
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x > gpurun_out/j6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j6_pytest.log
tail -25 gpurun_out/j6_pytest.log
timeout 300 python tools/train_bench.py > gpurun_out/j6_train_bench.log 2>&1; tail -3 gpurun_out/j6_train_bench.log
timeout 300 python tools/mc_bench.py --lego --reps 4 > gpurun_out/j6_mc_lego.log 2>&1; tail -5 gpurun_out/j6_mc_lego.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/j6_train_launches.csv python tools/train_profile.py > gpurun_out/j6_train_ncu.log 2>&1
