#!/bin/bash
# round-2 GPU job 26 (4 GPUs): final build — sharded bit-identity at 2 and 4 ranks, bench line N = 4
set -x
mkdir -p gpurun_out
python -m nerfmeshes_b200.build > gpurun_out/j26_build.log 2>&1 || { tail -5 gpurun_out/j26_build.log; exit 9; }
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/j26_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j26_pytest.log
tail -3 gpurun_out/j26_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 4 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/j26_bench_n4.json 2> gpurun_out/j26_bench_n4.err; echo "bench N=4 rc=$?"
tail -c 200 gpurun_out/j26_bench_n4.json
