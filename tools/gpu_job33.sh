#!/bin/bash
# round-2 GPU job 33 (1 GPU): the round's last build — full GPU suite + smoke
set -x
mkdir -p gpurun_out
python -m nerfmeshes_b200.build > gpurun_out/j33_build.log 2>&1 || { tail -5 gpurun_out/j33_build.log; exit 9; }
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/j33_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j33_pytest.log
tail -4 gpurun_out/j33_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/j33_smoke.log 2>&1; tail -2 gpurun_out/j33_smoke.log
