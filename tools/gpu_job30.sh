#!/bin/bash
# round-2 GPU job 30 (1 GPU): fused compositor with 16 / 32 rays per tile
set -x
mkdir -p gpurun_out
python -m nerfmeshes_b200.build > gpurun_out/j30_build.log 2>&1 || { tail -5 gpurun_out/j30_build.log; exit 9; }
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fused_compositor or edge_cases" > gpurun_out/j30_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j30_pytest.log
tail -12 gpurun_out/j30_pytest.log
