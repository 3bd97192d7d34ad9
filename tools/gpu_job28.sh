#!/bin/bash
# round-2 GPU job 28 (1 GPU): mbarrier waits with a long suspend-time hint vs plain polling, same box
set -x
mkdir -p gpurun_out
python -m nerfmeshes_b200.build > gpurun_out/j28_build.log 2>&1 || { tail -5 gpurun_out/j28_build.log; exit 9; }
timeout 600 python bench.py --only --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/j28_bench_hint.json 2> gpurun_out/j28_bench_hint.err; echo "bench rc=$?"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train.py -m gpu -q -x > gpurun_out/j28_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j28_pytest.log
tail -3 gpurun_out/j28_pytest.log
timeout 300 python tools/train_bench.py > gpurun_out/j28_train_bench_hint.log 2>&1; tail -1 gpurun_out/j28_train_bench_hint.log
NM_NVCC_EXTRA="-DNM_WAIT_HINT_NS=0" python -m nerfmeshes_b200.build --force > gpurun_out/j28_build_poll.log 2>&1 || { tail -5 gpurun_out/j28_build_poll.log; exit 9; }
timeout 600 python bench.py --only --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/j28_bench_poll.json 2> gpurun_out/j28_bench_poll.err; echo "bench poll rc=$?"
timeout 300 python tools/train_bench.py > gpurun_out/j28_train_bench_poll.log 2>&1; tail -1 gpurun_out/j28_train_bench_poll.log
timeout 600 python bench.py --only --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/j28_bench_poll2.json 2> gpurun_out/j28_bench_poll2.err
