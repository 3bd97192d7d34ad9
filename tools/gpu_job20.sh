#!/bin/bash
# round-2 GPU job 20 (1 GPU): mask prefetch in the chain; sensitivity of the training kernels to the weight-ring depth
set -x
mkdir -p gpurun_out
python -m nerfmeshes_b200.build > gpurun_out/j20_build.log 2>&1 || { tail -5 gpurun_out/j20_build.log; exit 9; }
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q > gpurun_out/j20_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j20_pytest.log
tail -3 gpurun_out/j20_pytest.log
for i in 1 2; do timeout 300 python tools/train_bench.py > gpurun_out/j20_train_bench_$i.log 2>&1; tail -1 gpurun_out/j20_train_bench_$i.log; done
NM_TC_STAGES=3 timeout 300 python tools/train_bench.py > gpurun_out/j20_train_bench_s3.log 2>&1; tail -1 gpurun_out/j20_train_bench_s3.log
NM_TC_STAGES=4 timeout 300 python tools/train_bench.py > gpurun_out/j20_train_bench_s4.log 2>&1; tail -1 gpurun_out/j20_train_bench_s4.log
NM_TRAIN_ACT_MN=1 timeout 300 python tools/train_bench.py > gpurun_out/j20_train_bench_actmn.log 2>&1; tail -1 gpurun_out/j20_train_bench_actmn.log
NM_TRAIN_ACT_MN=1 timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x > gpurun_out/j20_pytest_actmn.log 2>&1; tail -3 gpurun_out/j20_pytest_actmn.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/j20_train_launches.csv python tools/train_profile.py 4096 > gpurun_out/j20_train_ncu.log 2>&1
