#!/bin/bash
# round-2 GPU job 11 (1 GPU): float4 head kernel, fused encode->pack, warp-per-ray compositor adjoint; fp32 per-layer grads; power trace
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q > gpurun_out/j11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j11_pytest.log
tail -12 gpurun_out/j11_pytest.log
timeout 300 python tools/train_bench.py > gpurun_out/j11_train_bench.log 2>&1; tail -2 gpurun_out/j11_train_bench.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/j11_train_launches.csv python tools/train_profile.py 4096 > gpurun_out/j11_train_ncu.log 2>&1
