"""Multi-GPU equivalence as a driver-run test: the volume sharded over 2 GPUs (one process per GPU, torchrun) must reproduce the 1-GPU
tracker bit for bit -- poses, model maps, volume slabs, shift events, slices as multisets (tools/mgpu_check.py).  Skips itself on a
box with fewer than 2 GPUs."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("odometry", [0, 2])
def test_two_gpu_sharded_volume_equals_single_gpu(built, odometry):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29520 + odometry
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tools", "mgpu_check.py"), "--vol", "256", "--frames", "8", "--voxel-shift", "2", "--odometry", str(odometry)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, PYTHONPATH=ROOT))
    line = [l for l in r.stdout.splitlines() if l.startswith("MGPU_CHECK")]
    assert r.returncode == 0 and line, r.stdout[-3000:] + r.stderr[-3000:]
    assert "poses_equal=True" in line[0] and "slabs_equal=True" in line[0] and "model_maps_equal=True" in line[0] and "slices_equal=True" in line[0], line[0]
