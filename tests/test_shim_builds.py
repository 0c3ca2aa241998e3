"""The C++ shim (include/kintinuous_b200_shim.hpp) compiles and links against the C-ABI library: a caller written against the
reference's operator names builds unchanged."""
import os
import subprocess

from conftest import ROOT


def test_shim_compiles_and_links(built, tmp_path):
    exe = str(tmp_path / "shim_test")
    cmd = ["g++", "-std=c++14", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include", os.path.join(ROOT, "tests", "cpp", "shim_compile_test.cpp"),
           "-L", os.path.join(ROOT, "kintinuous_b200"), "-lkintinuous_b200", "-L", "/usr/local/cuda/lib64", "-lcudart",
           "-Wl,-rpath," + os.path.join(ROOT, "kintinuous_b200"), "-o", exe]
    subprocess.check_call(cmd)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "shim" in out.stdout
