"""bench.py prints ONE JSON line on stdout with the keys the driver reads (CPU: the reference arm; GPU: our arm)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e")


def run_bench(*args):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True,
                       cwd=ROOT, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must hold exactly one line, got %d" % len(lines)
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--envs", "512")
    for k in BASE_KEYS + ("impl", "cpu_baseline"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "shielded agent-steps/s" and d["unit"] == "agent-steps/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


@pytest.mark.gpu
def test_our_arm_line():
    d = run_bench("--envs", "8192", "--steps", "3", "--warmup", "3", "--cpu-seconds", "1")
    for k in BASE_KEYS + ("clocks", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["dtype"] == "f64" and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["gpu_launches"] >= 3
    assert d["e2e"]["h2d_bytes_per_step"] == 8192 * 12 and d["e2e"]["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["value"] > 0 and c["cores"] >= 1 and "sample" in c
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
