"""The C++ shim (include/kintinuous_b200_shim.hpp): a caller written against the reference's operator names and container types
(cuda/internal.h:299-536, containers/*.hpp, TSDFVolume.h, ColorVolume.h) builds unchanged against it and links against the C-ABI library.
The build check runs everywhere; the run check (the operators actually execute on the device through the shim) needs a GPU."""
import os
import subprocess

import pytest

from conftest import ROOT


def _build(tmp_path):
    exe = str(tmp_path / "shim_test")
    cmd = ["g++", "-std=c++14", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include", os.path.join(ROOT, "tests", "cpp", "shim_compile_test.cpp"),
           "-L", os.path.join(ROOT, "kintinuous_b200"), "-lkintinuous_b200", "-L", "/usr/local/cuda/lib64", "-lcudart",
           "-Wl,-rpath," + os.path.join(ROOT, "kintinuous_b200"), "-o", exe]
    subprocess.check_call(cmd)
    return exe


def test_shim_compiles_and_links(built, tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "shim" in out.stdout


@pytest.mark.gpu
def test_shim_runs_operators_on_the_device(built, tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "shim ok: 120 x 160 -> 60 x 80" in out.stdout, out.stdout
    assert "volume ok" in out.stdout and "262144 voxels" in out.stdout, out.stdout
    assert "checks ok" in out.stdout, out.stdout


def _build_facade(tmp_path):
    exe = str(tmp_path / "facade_test")
    cmd = ["g++", "-std=c++14", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include", os.path.join(ROOT, "tests", "cpp", "tracker_facade_test.cpp"),
           "-L", os.path.join(ROOT, "kintinuous_b200"), "-lkintinuous_b200", "-L", "/usr/local/cuda/lib64", "-lcudart", "-lpthread",
           "-Wl,-rpath," + os.path.join(ROOT, "kintinuous_b200"), "-o", exe]
    subprocess.check_call(cmd)
    return exe


def test_tracker_facade_compiles_and_links(built, tmp_path):
    """include/kintinuous_b200_tracker.hpp: a caller written against KintinuousTracker / OdometryProvider / CloudSlice / Resolution / Volume
    (the per-frame calls of backend/TrackerInterface.cpp:82-104) builds against the facade."""
    exe = _build_facade(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


@pytest.mark.gpu
def test_tracker_facade_runs_on_the_device(built, tmp_path):
    """...and runs: 15 frames through the 10-argument processFrame, +x shifts handed over under cloudMutex, live image / live TSDF taps,
    finalise, and ICPOdometry used on its own through the OdometryProvider interface."""
    exe = _build_facade(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "facade checks ok" in out.stdout, out.stdout + out.stderr
