#!/bin/bash
# round-2 GPU job 25 (2 GPUs): final build under torchrun — sharded bit-identity at 2 ranks, bench line N = 2
set -x
mkdir -p gpurun_out
python -m nerfmeshes_b200.build > gpurun_out/j25_build.log 2>&1 || { tail -5 gpurun_out/j25_build.log; exit 9; }
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/j25_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j25_pytest.log
tail -3 gpurun_out/j25_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/j25_bench_n2.json 2> gpurun_out/j25_bench_n2.err; echo "bench N=2 rc=$?"
tail -c 200 gpurun_out/j25_bench_n2.json
