#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_marching_cubes.py tests/test_gpu_multi.py tests/test_gpu_wide_parity.py tests/test_compat_train_eval.py -m gpu -q -s > gpurun_out/j4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j4_pytest.log
grep -n "pose [0-9]\|passed\|failed\|Error" gpurun_out/j4_pytest.log | tail -40
python bench.py --workload mesh --only --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/j4_mesh1.json 2> gpurun_out/j4_mesh1.err; echo "mesh1 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload mesh --only --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/j4_mesh2.json 2> gpurun_out/j4_mesh2.err; echo "mesh2 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/j4_mesh_launches.csv \
  python bench.py --workload mesh --only --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/j4_mesh_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mc_sign_kernel -s 1 -c 1 -o gpurun_out/j4_mc_sign \
  python bench.py --workload mesh --only --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/j4_sign_ncu.log 2>&1
