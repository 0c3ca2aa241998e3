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


def test_two_gpu_split_icp_allreduce_in_kernel(built):
    """KT_MG_SPLIT_ICP: the pixel rows of every ICP level are split over the ranks and the 29 normal-equation sums are all-reduced INSIDE
    icp_frame_kernel over NVLink peer memory (grid_sum_words_mg: one system-scope red.add per rank and component, local poll) -- the
    north_star's per-iteration all-reduce without a collective library.  Every rank must hold bit-identical poses (integer sums); against
    the 1-GPU run the poses agree to rounding (<= 1e-5: the float partial sums are grouped differently), shift events are identical and
    the fused TSDF is within 1 LSB on >= 99.9 % of the voxels."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29527",
           os.path.join(ROOT, "tools", "mgpu_check.py"), "--vol", "256", "--frames", "8", "--voxel-shift", "2", "--odometry", "0", "--split-icp"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, PYTHONPATH=ROOT))
    line = [l for l in r.stdout.splitlines() if l.startswith("MGPU_SPLIT_CHECK")]
    assert r.returncode == 0 and line, r.stdout[-3000:] + r.stderr[-3000:]
    assert "poses_within_1e-5=True" in line[0] and "ranks_identical=True" in line[0], line[0]
