"""bench.py's reference arm runs on the host cores (the CPU restatement of the reference's path): its JSON line can be
checked in the CPU-only container.  The GPU arm is exercised by the driver and by `-m gpu` below."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config"}


def run(*args):
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), cwd=ROOT, text=True,
                                  stderr=subprocess.DEVNULL)
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, "bench.py must print exactly one line on stdout"
    return json.loads(lines[0])


def test_reference_arm_json_line():
    d = run("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert KEYS <= set(d) and d["impl"] == "reference"
    assert d["unit"] == "env-steps/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["steps"] == 1 and d["warmup"] == 1 and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.gpu
def test_gpu_arm_json_line():
    d = run("--steps", "5", "--warmup", "3", "--envs", "32768")
    assert KEYS <= set(d) and d.get("impl") != "reference"
    assert d["value"] > 0 and d["gpu_launches"] == 5 and d["steps"] == 5
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and 0 < r["frac"] < 1 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
