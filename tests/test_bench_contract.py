"""bench.py prints exactly one JSON line with the keys the driver relies on (both arms)."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def _run(*args, timeout=600):
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), *args], capture_output=True, text=True,
                         timeout=timeout, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.strip().splitlines()          # nothing but the JSON line may reach stdout (library banners go to stderr)
    assert len(lines) == 1 and lines[0].startswith("{"), out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line_small_workload():
    """--impl reference runs without a GPU (oracle port on the host cores)."""
    d = _run("--impl", "reference", "--workload", "small", "--steps", "2", "--warmup", "1")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["unit"] == "queries/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"]


@pytest.mark.gpu
def test_b200_arm_line_small_workload():
    d = _run("--workload", "small", "--steps", "3", "--warmup", "3", "--cpu-sample", "8")
    assert BASE_KEYS | {"roofline", "clocks"} <= set(d)
    assert d["n_gpus"] == 1 and d["value"] > 0 and d["gpu_launches"] > 0 and d["data"] == "synthetic"
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 100_000_000 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]                      # the host-buffer path includes the copies
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["value"] > 0 and c["parity_on_sample"]["topk_rows_identical"] == 1.0
    assert c["parity_on_sample"]["vote_identical"] == 1.0
