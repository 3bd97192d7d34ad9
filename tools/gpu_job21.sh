#!/bin/bash
# round-2 GPU job 21 (1 GPU): forward kernel vs weight-ring depth
set -x
mkdir -p gpurun_out
python -m nerfmeshes_b200.build > gpurun_out/j21_build.log 2>&1 || { tail -5 gpurun_out/j21_build.log; exit 9; }
for s in 7 6 5 4; do
  NM_TC_STAGES=$s timeout 600 python bench.py --only --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/j21_bench_s$s.json 2> gpurun_out/j21_bench_s$s.err; echo "s=$s rc=$?"
done
timeout 300 python tools/mlp_bench.py > gpurun_out/j21_mlp_bench.log 2>&1; tail -3 gpurun_out/j21_mlp_bench.log
