#!/bin/bash
# round-2 GPU job 2 (2 GPUs): marching cubes parity, multi-GPU bit-identity tests, row/slab-sharded bench, MC launch list
set -x
mkdir -p gpurun_out
python -m pytest tests/test_marching_cubes.py tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/j2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j2_pytest.log
tail -15 gpurun_out/j2_pytest.log
python bench.py --workload mesh --only --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/j2_mesh1.json 2> gpurun_out/j2_mesh1.err; echo "mesh1 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload mesh --only --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/j2_mesh2.json 2> gpurun_out/j2_mesh2.err; echo "mesh2 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/j2_bench2.json 2> gpurun_out/j2_bench2.err; echo "bench2 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/j2_mesh_launches.csv \
  python bench.py --workload mesh --only --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/j2_mesh_ncu.log 2>&1
tail -c 600 gpurun_out/j2_mesh1.json; tail -3 gpurun_out/j2_mesh1.err gpurun_out/j2_mesh2.err gpurun_out/j2_bench2.err
