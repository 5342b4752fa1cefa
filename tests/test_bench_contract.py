"""bench.py contract: one JSON line with the keys the driver reads.  The reference arm runs on CPU (bounded sample);
the GPU arm is checked under -m gpu with a small batch."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "gpu_launches"}


def _run(args, timeout):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, f"bench.py must print exactly one line, got {len(lines)}"
    return json.loads(lines[0])


def test_reference_arm_prints_one_json_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-workers", "4"], 600)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 4
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("Model.detect") and d["vs_baseline"] is None


@pytest.mark.gpu
def test_gpu_arm_prints_one_json_line():
    d = _run(["--steps", "2", "--warmup", "3", "--batch", "20", "--no-cpu-baseline"], 900)
    assert BASE_KEYS | {"roofline", "roofline_pyramid", "roofline_cascade", "clocks"} <= set(d)
    assert d["n_gpus"] == 1 and d["scaling"] == "weak" and d["value"] > 0 and d["gpu_launches"] > 0
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 20 * 1080 * 1920 and d["e2e"]["d2h_bytes_per_step"] > 0
    for k in ("roofline", "roofline_pyramid", "roofline_cascade"):
        r = d[k]
        assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0 < r["frac"] < 1.5 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["config"]["windows_per_frame"] == 3045278 and d["config"]["levels"] == 64
