#!/bin/bash
# round-2 GPU job 23 (1 GPU): CTA pairs sharing one multicast weight stream — parity suite, bench with and without
set -x
mkdir -p gpurun_out
python -m nerfmeshes_b200.build > gpurun_out/j23_build.log 2>&1 || { tail -5 gpurun_out/j23_build.log; exit 9; }
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/j23_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/j23_parity.log
tail -5 gpurun_out/j23_parity.log
timeout 600 python bench.py --only --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/j23_bench.json 2> gpurun_out/j23_bench.err; echo "bench rc=$?"
NM_TC_CLUSTER=0 timeout 600 python bench.py --only --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/j23_bench_nocluster.json 2> gpurun_out/j23_bench_nocluster.err; echo "bench nocluster rc=$?"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/j23_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j23_pytest.log
tail -5 gpurun_out/j23_pytest.log
timeout 300 python tools/train_bench.py > gpurun_out/j23_train_bench.log 2>&1; tail -1 gpurun_out/j23_train_bench.log
