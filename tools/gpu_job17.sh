#!/bin/bash
# round-2 GPU job 17 (8 GPUs): sharded render / mesh bit-identity at 2, 4, 8 ranks; bench line at N = 8
set -x
mkdir -p gpurun_out
python -m nerfmeshes_b200.build > gpurun_out/j17_build.log 2>&1 || { tail -5 gpurun_out/j17_build.log; exit 9; }
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "8" > gpurun_out/j17_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j17_pytest.log
tail -4 gpurun_out/j17_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/j17_bench_n8.json 2> gpurun_out/j17_bench_n8.err; echo "bench N=8 rc=$?"
tail -c 300 gpurun_out/j17_bench_n8.json
