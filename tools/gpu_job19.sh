#!/bin/bash
# round-2 GPU job 19 (1 GPU): dZ packs as MN-major tiles through per-warp bulk stores (mode 2)
set -x
mkdir -p gpurun_out
python -m nerfmeshes_b200.build > gpurun_out/j19_build.log 2>&1 || { tail -5 gpurun_out/j19_build.log; exit 9; }
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q > gpurun_out/j19_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j19_pytest.log
tail -6 gpurun_out/j19_pytest.log
timeout 300 python tools/train_bench.py > gpurun_out/j19_train_bench.log 2>&1; tail -1 gpurun_out/j19_train_bench.log
NM_TRAIN_DZ_MN=0 timeout 300 python tools/train_bench.py > gpurun_out/j19_train_bench_k.log 2>&1; tail -1 gpurun_out/j19_train_bench_k.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/j19_train_launches.csv python tools/train_profile.py 4096 > gpurun_out/j19_train_ncu.log 2>&1
