#!/bin/bash
# round-2 GPU job 3 (2 GPUs): the whole GPU suite incl. multi-GPU bit-identity, mesh bench at N=1/2, MC launch list + sign-pass capture
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/j3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j3_pytest.log
tail -25 gpurun_out/j3_pytest.log
python bench.py --workload mesh --only --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/j3_mesh1.json 2> gpurun_out/j3_mesh1.err; echo "mesh1 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload mesh --only --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/j3_mesh2.json 2> gpurun_out/j3_mesh2.err; echo "mesh2 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/j3_mesh_launches.csv \
  python bench.py --workload mesh --only --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/j3_mesh_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mc_sign_kernel -s 1 -c 1 -o gpurun_out/j3_mc_sign \
  python bench.py --workload mesh --only --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/j3_sign_ncu.log 2>&1
tail -c 400 gpurun_out/j3_mesh1.json
