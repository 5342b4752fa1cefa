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
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-workers", "4", "--cpu-kind", "port", "--no-single-thread"], 600)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 4
    assert d["native_so_loaded"] == [], "the CPU arm must not map the product library"
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


@pytest.mark.gpu
@pytest.mark.parametrize("config", ["A", "D"])
def test_small_configs_print_one_json_line_and_match_their_golden(config):
    """bench.py --config A / D: BASELINE configs[0] and configs[3], each compared with the reference's own output
    (tests/golden/) inside the run."""
    d = _run(["--config", config, "--steps", "2", "--warmup", "3"], 900)
    assert (BASE_KEYS - {"cpu_baseline"}) <= set(d) and d["value"] > 0 and d["unit"] == "frames/s"
    assert d["config"]["matches_reference_golden"] is True and d["config"]["eval_cost"] > 1


def test_cpu_arm_inputs_match_the_product_side():
    """oracle/cpu_arm.py restates the frame recipe and parses the model file on its own (it never imports the product):
    both must give what the GPU arm uses."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_arm
    import waldboost_b200 as wb
    from waldboost_b200 import synthetic as S
    for seed, h, w in ((1000, 1080, 1920), (1003, 200, 260), (5, 64, 96)):
        assert np.array_equal(cpu_arm.synthetic_frame(seed, h, w), S.synthetic_frame(seed, h, w))
    path = os.path.join(ROOT, "tests", "golden", "configB_model.pb")
    d, M = cpu_arm.read_model_pb(path), wb.Model.load(path)
    assert tuple(d["shape"]) == tuple(M.shape) and len(d["trees"]) == len(M) == 1024
    assert np.array_equal(np.array(d["theta"], np.float32), np.array(M.theta, np.float32))
    for (f, t, l, r, p), w in list(zip(d["trees"], M.classifier))[::97]:
        assert np.array_equal(f, np.asarray(w.feature).reshape(-1, 3)) and np.array_equal(t, w.threshold)
        assert np.array_equal(l, w.left) and np.array_equal(r, w.right) and np.array_equal(p, w.prediction)
    assert d["func"].endswith("grad_hist") and (d["shrink"], d["n_per_oct"], d["smooth"]) == (2, 8, 1)
    shards = cpu_arm.level_shards(d, 1080, 1920, 4)
    assert sorted(sum(shards, [])) == list(range(64))
