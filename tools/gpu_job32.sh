#!/bin/bash
# round-2 GPU job 32 (1 GPU): front-end warps emit half of the training forward's packs (NM_TRAIN_FE_EMIT=1)
set -x
mkdir -p gpurun_out
python -m nerfmeshes_b200.build > gpurun_out/j32_build.log 2>&1 || { tail -5 gpurun_out/j32_build.log; exit 9; }
NM_TRAIN_FE_EMIT=1 timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x > gpurun_out/j32_pytest_fe.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j32_pytest_fe.log
tail -4 gpurun_out/j32_pytest_fe.log
NM_TRAIN_FE_EMIT=1 timeout 300 python tools/train_bench.py > gpurun_out/j32_train_bench_fe.log 2>&1; tail -1 gpurun_out/j32_train_bench_fe.log
timeout 300 python tools/train_bench.py > gpurun_out/j32_train_bench.log 2>&1; tail -1 gpurun_out/j32_train_bench.log
NM_TRAIN_FE_EMIT=1 timeout 300 python tools/train_bench.py > gpurun_out/j32_train_bench_fe2.log 2>&1; tail -1 gpurun_out/j32_train_bench_fe2.log
