#!/bin/bash
# round-2 GPU job 12 (1 GPU): full GPU suite, default bench + reference arm, launch lists, full ncu captures of the training kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/j12_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j12_pytest.log
tail -8 gpurun_out/j12_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/j12_smoke.log 2>&1; tail -2 gpurun_out/j12_smoke.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/j12_bench.json 2> gpurun_out/j12_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/j12_bench_ref.json 2> gpurun_out/j12_bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --workload buff --only --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/j12_buff.json 2> gpurun_out/j12_buff.err
timeout 600 python bench.py --workload fern --only --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/j12_fern.json 2> gpurun_out/j12_fern.err
timeout 600 python bench.py --workload mesh --only --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/j12_mesh.json 2> gpurun_out/j12_mesh.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/j12_bench_launches.csv \
  python bench.py --only --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/j12_bench_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/j12_buff_launches.csv \
  python bench.py --workload buff --only --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/j12_buff_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/j12_train_launches.csv python tools/train_profile.py 4096 > gpurun_out/j12_train_ncu.log 2>&1
# full captures: the 2nd step's fused kernels (2 x mode 1, 2 x mode 2) and the widest weight-gradient GEMMs
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 4 -c 4 -o gpurun_out/j12_train_mlp python tools/train_profile.py 4096 > gpurun_out/j12_ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 24 -c 6 -o gpurun_out/j12_train_gemm python tools/train_profile.py 4096 > gpurun_out/j12_ncu_b.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"invcdf_kernel|aabb_kernel|head_backward" -c 3 -o gpurun_out/j12_light python bench.py --workload buff --only --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/j12_ncu_c.log 2>&1
timeout 300 python tools/parity_report.py > gpurun_out/j12_parity_report.json 2> gpurun_out/j12_parity.err
timeout 300 python tools/train_bench.py > gpurun_out/j12_train_bench.log 2>&1; tail -1 gpurun_out/j12_train_bench.log
tail -c 400 gpurun_out/j12_bench.json
