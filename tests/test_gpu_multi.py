"""Multi-GPU parity: one process per GPU under torchrun + NCCL; see tests/_multi_worker.py for the checks
(row-sharded image bit-identical to the single-GPU image for lego / fern-NDC / BuFF; slab-sharded mesh equal to the
single-GPU mesh array for array).  Skipped on boxes with fewer than 2 GPUs."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_render_and_mesh_match_single_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + (os.getpid() + world) % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_multi_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and f"MULTI_OK {world}" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
