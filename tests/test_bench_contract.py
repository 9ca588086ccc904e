"""bench.py contract checks that run without a GPU: the reference arm (`--impl reference`: the
unmodified reference on the host cores) prints exactly ONE JSON line on stdout with the contract's
keys, rank > 0 prints nothing, and the GPU arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

from _checkers import Reference

ROOT = Path(__file__).resolve().parent.parent


def _bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True,
                          timeout=900, env=e)


@pytest.mark.skipif(not Reference.available(), reason="oracle/_ref/libfastc_ref.so not built")
def test_reference_arm_prints_one_contract_line():
    r = _bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mpix/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("BC7 Mpix/s at -q 50")
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    r = _bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
               env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _bench("--steps", "1", "--warmup", "3")
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CPU fallback" in r.stderr
