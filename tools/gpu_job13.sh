#!/bin/bash
# round-2 GPU job 13 (4 GPUs): sharded render / mesh bit-identity at 2 and 4 ranks, bench lines at N = 2 and 4
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/j13_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j13_pytest.log
tail -4 gpurun_out/j13_pytest.log
for N in 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/j13_bench_n$N.json 2> gpurun_out/j13_bench_n$N.err; echo "bench N=$N rc=$?"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --impl reference --steps 2 --warmup 1 > gpurun_out/j13_ref_n4.json 2> gpurun_out/j13_ref_n4.err; echo "ref rc=$?"
tail -c 300 gpurun_out/j13_bench_n4.json
