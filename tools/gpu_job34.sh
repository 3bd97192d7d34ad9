#!/bin/bash
# round-2 GPU job 34 (1 GPU): gradient tests three times (run-to-run atomic-order noise), then the full suite
set -x
mkdir -p gpurun_out
python -m nerfmeshes_b200.build > gpurun_out/j34_build.log 2>&1 || { tail -5 gpurun_out/j34_build.log; exit 9; }
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q > gpurun_out/j34_train_$i.log 2>&1; tail -1 gpurun_out/j34_train_$i.log; done
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/j34_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j34_pytest.log
tail -3 gpurun_out/j34_pytest.log
