"""Timings of the BASELINE.json configurations that are not the bench headline (A, C, D) on one B200, with random-init
cascades of the stated shape and 'wald' thresholds calibrated on the GPU path itself (trace_windows on a window sample).
Prints one JSON object per configuration.  Usage: python profiles/run_configs.py [A C D]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import waldboost_b200 as wb
from waldboost_b200 import synthetic as S
from waldboost_b200.engine import get_engine


def calibrated_model(shape, opts, n_stages, depth, frame, keep_total, n_sample=60000, seed=7):
    """random trees with thresholds from the channel quantiles of `frame`; thetas from calibrate_thetas driven by the
    GPU's own per-stage leaf predictions on a random sample of windows of level 0."""
    lv0 = next(iter(wb.channels.channel_pyramid(frame, opts)))[0]
    lo, hi = S.channel_quantiles(lv0)
    trees = S.random_trees(shape, n_stages, depth, lo, hi, seed=seed)
    M = wb.Model(shape, opts)
    for t in trees:
        M.append(t, -np.inf)
    rng = np.random.default_rng(1)
    u, v, _ = lv0.shape
    rs = rng.integers(0, u - shape[0], n_sample)
    cs = rng.integers(0, v - shape[1], n_sample)
    leaf, _ = M.trace_windows(lv0, rs, cs)
    pred = np.stack([t.prediction[leaf[:, k]] for k, t in enumerate(trees)], axis=1)

    def stage(t, alive):
        return pred[:, t] if alive is None else pred[alive, t]
    M.theta = [float(x) for x in S.calibrate_thetas(stage, n_stages, keep_total)]
    return M


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


def config_A():
    opts = dict(shrink=2, n_per_oct=8, smooth=1, channels=wb.channels.grad_hist)
    M = wb.Model.load(os.path.join(ROOT, "tests", "golden", "configA_model.pb"))
    frame = S.synthetic_frame(1000, 480, 640)
    dt, boxes = timed(lambda: M.detect(frame), 20)
    return {"config": "A: Model.detect, one 640x480 uint8 frame, 12x12x4, 256 depth-2 stages (reference-calibrated model)",
            "ms_per_frame": dt * 1e3, "hits": len(boxes), "eval_cost": M.eval_cost}


def config_C():
    opts = dict(shrink=2, n_per_oct=8, smooth=1, channels=wb.channels.grad_mag_hist)
    frame = S.synthetic_frame(1000, 2160, 3840)
    M = calibrated_model((20, 20, 10), opts, 256, 2, S.synthetic_frame(1000, 540, 960), 1e-4)
    M.reset()
    dt, boxes = timed(lambda: M.detect(frame), 3)
    eng = get_engine()
    plan = M._plan(eng, 2160, 3840)
    dev = eng.upload_images(frame[None])
    tp, _ = timed(lambda: eng.pyramid(dev, plan), 3)
    return {"config": "C: Model.detect, one 3840x2160 uint8 frame, 20x20x10 grad_mag+grad_hist(9), 256 depth-2 stages, one GPU",
            "ms_per_frame": dt * 1e3, "pyramid_ms": tp * 1e3, "pyramid_GBps": (2160 * 3840 + 4 * plan.chn_floats) / tp / 1e9,
            "hits": len(boxes), "eval_cost": M.eval_cost, "windows": int(plan.info.n_loc)}


def config_D(n_images=1000):
    opts = dict(shrink=2, n_per_oct=8, smooth=1, channels=wb.channels.grad_hist)
    frames = S.synthetic_frames(16, 480, 640)
    M = calibrated_model((12, 12, 4), opts, 2048, 4, frames[0], 1e-4)
    batch = np.ascontiguousarray(frames[np.arange(n_images) % 16])
    M.reset()
    dt, (out, hits) = timed(lambda: M.detect_batch(batch, return_hits=True), 2)
    return {"config": f"D: dense scoring (scan) of {n_images} 640x480 images, 2048-stage depth-4 cascade (complete depth-4 stage records)",
            "images_per_s": n_images / dt, "ms_per_image": dt * 1e3 / n_images, "hits_per_image": hits.size / n_images,
            "eval_cost": M.eval_cost}


if __name__ == "__main__":
    which = sys.argv[1:] or ["A", "C", "D"]
    for c in which:
        print(json.dumps({"A": config_A, "C": config_C, "D": config_D}[c]()), flush=True)
