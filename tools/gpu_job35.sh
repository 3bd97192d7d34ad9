#!/bin/bash
# round-2 GPU job 35 (1 GPU): parity tests with the tightened end-to-end bars
set -x
mkdir -p gpurun_out
python -m nerfmeshes_b200.build > gpurun_out/j35_build.log 2>&1 || { tail -5 gpurun_out/j35_build.log; exit 9; }
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wide_parity.py -m gpu -q > gpurun_out/j35_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j35_pytest.log
tail -5 gpurun_out/j35_pytest.log
