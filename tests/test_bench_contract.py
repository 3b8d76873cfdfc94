"""The bench contract that can be checked without a GPU: the reference arm's JSON line and the refusal to run the GPU arm
on a machine without CUDA (there is no CPU fallback to time)."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = _run("--impl", "reference", "--steps", "2", "--warmup", "1", "--ncol", "20000", "--model-steps", "2",
             "--cpu-sample-cols", "4000")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cell_timesteps_per_sec" and d["unit"] == "cell-steps/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["config"]["workload"] == "conus1km" and d["config"]["sampled_columns"] == 4000
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine WITHOUT a GPU")
def test_gpu_arm_refuses_to_run_without_cuda():
    r = _run("--steps", "1", "--warmup", "0", "--ncol", "1000", "--model-steps", "1", "--no-cpu-baseline", "--no-e2e")
    assert r.returncode != 0 and not r.stdout.strip()
    assert "CUDA" in r.stderr or "cuda" in r.stderr
