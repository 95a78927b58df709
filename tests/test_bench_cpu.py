"""bench.py contract checks that need no GPU: the reference arm's JSON line, and that the product refuses to run without a GPU / without
its CUDA library (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600,
                          env={**os.environ, **(env or {})})


def test_reference_arm_json_contract():
    r = _run("--impl", "reference", "--workload", "tiny", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "images_per_sec" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_runs_on_rank0_only():
    r = _run("--impl", "reference", "--workload", "tiny", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_respects_the_wall_budget():
    r = _run("--impl", "reference", "--workload", "tiny", "--steps", "3", "--warmup", "1", env={"CID_CPU_BUDGET_S": "0.01"})
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["cpu_baseline"]["iterations_timed"] == 1                      # cut short after the first measured iteration


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_product_arm_refuses_to_run_without_a_gpu():
    r = _run("--workload", "tiny", "--steps", "1", "--warmup", "0", "--no-cpu")
    assert r.returncode != 0 and "GPU" in (r.stderr + r.stdout)


def test_missing_library_fails_loudly(monkeypatch):
    from consistentid_b200 import lib
    monkeypatch.setattr(lib, "LIB_PATH", "/nonexistent/libcidb200.so")
    with pytest.raises(ImportError, match="no CPU/eager fallback"):
        lib._load()
