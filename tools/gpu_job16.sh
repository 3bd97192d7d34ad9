#!/bin/bash
# round-2 GPU job 16 (1 GPU): full GPU suite + smoke on the fused-compositor build
set -x
mkdir -p gpurun_out
python -m nerfmeshes_b200.build > gpurun_out/j16_build.log 2>&1 || { tail -5 gpurun_out/j16_build.log; exit 9; }
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/j16_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j16_pytest.log
tail -6 gpurun_out/j16_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/j16_smoke.log 2>&1; tail -2 gpurun_out/j16_smoke.log
timeout 300 python tools/parity_report.py > gpurun_out/j16_parity_report.json 2> gpurun_out/j16_parity.err; echo "parity rc=$?"
