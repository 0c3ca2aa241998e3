"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the contract's keys (here it falls back to
the CPU oracle port because there is no CUDA device), and the product arm refuses to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(args, env_extra=None, timeout=600):
    env = dict(os.environ, KT_BENCH_FRAMES="8", **(env_extra or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)


def test_reference_arm_prints_the_contract_line(built):
    r = _run(["--impl", "reference", "--steps", "2", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "frames/s" and d["vs_baseline"] is None
    for k in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("CUDA device present")
    r = _run(["--steps", "2", "--warmup", "1"])
    assert r.returncode != 0
    assert "no CUDA device" in (r.stdout + r.stderr)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]          # no number without the CUDA path
