"""Slab domain decomposition (BASELINE configs[4] family) on 2 GPUs: needs a 2-GPU box
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_dd.py -m gpu`); skipped on a single GPU."""
import os
import subprocess
import sys

import pytest

from msmpscu_b200 import capi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["thermal", "cascade"])
@pytest.mark.parametrize("world", [2, 4])
def test_slab_run_matches_single_gpu(world, mode):
    """world = 2: both neighbours of a rank are the same rank; world = 4: interior ranks with two distinct neighbours.
    cascade: PKA + electronic stopping + displacement-limited step (mdb_dd_run_sched) against mdb_run_sched"""
    if capi.load().mdb_device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29533 + world + (10 if mode == "cascade" else 0)), os.path.join(ROOT, "tests", "dd_worker.py"), "35", mode]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    print(out.stdout[-3000:], out.stderr[-3000:])
    assert out.returncode == 0
    assert "DD_RESULT PASS" in out.stdout
