#!/bin/bash
# round-2 GPU job 15 (1 GPU): fused compositor — bit-identity tests, full suite, bench
set -x
mkdir -p gpurun_out
python -m nerfmeshes_b200.build > gpurun_out/j15_build.log 2>&1 || { tail -5 gpurun_out/j15_build.log; exit 9; }
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fused_compositor" > gpurun_out/j15_fused.log 2>&1; echo "fused rc=$?" >> gpurun_out/j15_fused.log
tail -15 gpurun_out/j15_fused.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/j15_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j15_pytest.log
tail -6 gpurun_out/j15_pytest.log
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/j15_bench.json 2> gpurun_out/j15_bench.err; echo "bench rc=$?"
NM_FUSED_COMPOSITE=0 timeout 900 python bench.py --only --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/j15_bench_unfused.json 2> gpurun_out/j15_bench_unfused.err; echo "bench unfused rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/j15_bench_launches.csv \
  python bench.py --only --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/j15_bench_ncu.log 2>&1
tail -c 300 gpurun_out/j15_bench.json
