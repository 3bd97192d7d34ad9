#!/bin/bash
# round-2 GPU job 27 (1 GPU): last check of bench.py after its final edit
set -x
mkdir -p gpurun_out
python -m nerfmeshes_b200.build > gpurun_out/j27_build.log 2>&1 || { tail -5 gpurun_out/j27_build.log; exit 9; }
timeout 900 python bench.py > gpurun_out/j27_bench.json 2> gpurun_out/j27_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --workload mesh --only --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/j27_mesh.json 2> gpurun_out/j27_mesh.err; echo "mesh rc=$?"
tail -c 300 gpurun_out/j27_bench.json
