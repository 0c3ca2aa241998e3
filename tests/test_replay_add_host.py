"""kt_replay.cuh: replay_add(x, a, k) must return the bits of k sequential float additions x += a (the reference's running sums along z,
tsdf_volume.cu:565-574) -- it is what lets integrate_kernel start a column at any z without replaying the additions one by one.
Host build of the same header against the plain loop: random magnitudes, sign crossings, exact ties (|r| = ulp / 2), a far below the ulp
of x, x = 0, and the value ranges the kernel really uses (v_x = fx * p_x up to a few thousand, steps of 1e-3 .. 10)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(ROOT, "tests", "cpp", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libkt_replay_host.so")
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "kintinuous_b200", "csrc"), "-o", so,
                           os.path.join(ROOT, "tests", "cpp", "replay_host.cpp")])
    return C.CDLL(so)


def _check(lib, x, a, k):
    x = np.ascontiguousarray(x, np.float32); a = np.ascontiguousarray(a, np.float32); k = np.ascontiguousarray(k, np.int32)
    return lib.ktr_check_many(x.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p), k.ctypes.data_as(C.c_void_p), len(x))


def test_random_magnitudes_and_sign_crossings(lib):
    rng = np.random.default_rng(0)
    n = 600000
    x = rng.standard_normal(n) * 10.0 ** rng.integers(-3, 5, n)
    a = rng.standard_normal(n) * 10.0 ** rng.integers(-6, 2, n)
    k = rng.integers(0, 2049, n)
    assert _check(lib, x, a, k) == 0


def test_kernel_value_ranges(lib):
    rng = np.random.default_rng(1)
    n = 400000
    x = rng.uniform(-3500, 3500, n)                      # fx * p_x for a 6 m volume
    a = rng.uniform(-7, 7, n) * rng.choice([1.0, 0.1, 0.01], n)
    k = rng.integers(0, 2049, n)
    assert _check(lib, x, a, k) == 0


def test_ties_tiny_steps_and_zero(lib):
    rng = np.random.default_rng(2)
    xs = (np.float32(1.0) + rng.random(4000).astype(np.float32)) * np.float32(2.0) ** rng.integers(-4, 12, 4000).astype(np.float32)
    ulp = np.spacing(xs.astype(np.float32))
    at = ((rng.integers(0, 50, 4000) + 0.5) * ulp).astype(np.float32) * rng.choice([-1, 1], 4000).astype(np.float32)       # exact ties
    assert _check(lib, xs, at, rng.integers(0, 700, 4000)) == 0
    tiny = (ulp * rng.uniform(0.0, 0.6, 4000)).astype(np.float32) * rng.choice([-1, 1], 4000).astype(np.float32)             # around ulp / 2
    assert _check(lib, xs, tiny, rng.integers(0, 700, 4000)) == 0
    assert _check(lib, np.zeros(500), rng.standard_normal(500), rng.integers(0, 600, 500)) == 0
    assert _check(lib, rng.standard_normal(500), np.zeros(500), rng.integers(0, 600, 500)) == 0
    # powers of two and the values just around them (binade boundaries from both sides)
    p2 = np.float32(2.0) ** rng.integers(-6, 12, 6000).astype(np.float32)
    near = np.nextafter(p2, np.float32(0) if True else p2) 
    xb = np.concatenate([p2, near, np.nextafter(p2, np.float32(1e30))]) * np.concatenate([rng.choice([-1, 1], 18000)]).astype(np.float32)
    ab = (rng.standard_normal(18000) * 10.0 ** rng.integers(-5, 1, 18000)).astype(np.float32)
    assert _check(lib, xb, ab, rng.integers(0, 1500, 18000)) == 0


def test_stress_rounded_values(lib):
    """values with few significant bits (exact multiples, ties, signed zeros) mixed with generic ones"""
    rng = np.random.default_rng(123)
    n = 1000000
    x = rng.standard_normal(n) * 2.0 ** rng.integers(-12, 14, n)
    a = rng.standard_normal(n) * 2.0 ** rng.integers(-20, 8, n)
    a = np.where(rng.random(n) < 0.3, np.round(a * 2.0 ** rng.integers(0, 20, n)) / 2.0 ** rng.integers(0, 20, n), a)
    x = np.where(rng.random(n) < 0.2, np.round(x * 4) / 4, x)
    assert _check(lib, x, a, rng.integers(0, 1200, n)) == 0


def _check_fma(lib, x, m, f, k):
    x = np.ascontiguousarray(x, np.float32); m = np.ascontiguousarray(m, np.float32); f = np.ascontiguousarray(f, np.float32); k = np.ascontiguousarray(k, np.int32)
    return lib.ktr_check_many_fma(x.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p), f.ctypes.data_as(C.c_void_p), k.ctypes.data_as(C.c_void_p), len(x))


def test_replay_fma_matches_the_fused_loop(lib):
    """The step the reference build really executes along z is v = fma(m, f, v) (nvcc contracts v += R.z * cell * f, tsdf_volume.cu:574): the
    addend is the exact 48-bit product.  replay_fma(x, m, f, k) against k sequential fmaf: random magnitudes, rounded operands (exact
    products, ties), zeros, and the kernel's ranges (m = R.z * cell ~ 1e-6 .. 1e-2, f = fx ~ 528, x = fx * p_x up to a few thousand)."""
    rng = np.random.default_rng(5)
    n = 800000
    x = rng.standard_normal(n) * 2.0 ** rng.integers(-12, 14, n)
    m = rng.standard_normal(n) * 2.0 ** rng.integers(-14, 2, n)
    f = rng.standard_normal(n) * 2.0 ** rng.integers(-2, 11, n)
    m = np.where(rng.random(n) < 0.3, np.round(m * 2.0 ** rng.integers(0, 16, n)) / 2.0 ** rng.integers(0, 16, n), m)
    f = np.where(rng.random(n) < 0.3, np.round(f * 4) / 4, f)
    x = np.where(rng.random(n) < 0.2, np.round(x * 4) / 4, x)
    assert _check_fma(lib, x, m, f, rng.integers(0, 1200, n)) == 0
    n = 600000
    x = rng.uniform(-3500, 3500, n); m = rng.uniform(-1, 1, n) * 0.0117 * rng.choice([1, 0.1, 0.01, 1e-4], n)
    f = np.full(n, 528.0144) * rng.choice([1, 2, 0.5], n)
    assert _check_fma(lib, x, m, f, rng.integers(0, 2049, n)) == 0
    z = np.zeros(2000)
    assert _check_fma(lib, z, rng.standard_normal(2000), rng.standard_normal(2000), rng.integers(0, 50, 2000)) == 0
    assert _check_fma(lib, rng.standard_normal(2000), z, rng.standard_normal(2000), rng.integers(0, 50, 2000)) == 0
