"""Generate the golden fixtures in this directory by running the UNMODIFIED reference (/root/reference) under the
shims of oracle/ref_harness.py.  Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Outputs (all small, committed):
  small_pyramid.npz    96x128 uint8 + float32 frames, every level of channel_pyramid for 4 channel configurations
  small_model.pb       24-stage depth-2 cascade written by the reference's Model.save
  small_detect.npz     reference Model.detect / predict_on_image outputs for that model (dense + wald thetas)
  configA_model.pb     BASELINE config A cascade (256 depth-2 stages, wald thetas), written by the reference
  configA_detect.npz   reference Model.detect on the 640x480 config-A frame: boxes, scores, n_loc, n_weak,
                       per-level float64 channel sums
  generic_model.pb     unbalanced / depth-3 trees (generic-topology path) + reference outputs in generic_detect.npz
  configB_model.pb / configB_detect.npz   (--config-b)  the model bench.py times + reference detect() on a 540x960 crop
  fpga_pyramid.npz     (--fpga)      integer channels of the reference's FPGA variant
  multi_*.pb / multi_detect.npz      (--multi)     waldboost.detect(image, A, B, response_scale=...) of the reference:
                       boxes, scores, labels in the reference's order (__init__.py:75-130)
  configC_model.pb / configC_detect.npz  (--config-c)  one 3840x2160 frame, 20x20x10 grad_mag + grad_hist(9) channels,
                       256 depth-2 stages: the reference's detect() at BASELINE config C's stated size
  configD_model.pb / configD_scan.npz    (--config-d)  a 2048-stage depth-4 cascade through the reference's
                       Model.scan_channels (the dense-scoring call of Pool.update, BASELINE config D)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_harness  # noqa: E402

wb = ref_harness.import_reference()
from waldboost import channels as rch  # noqa: E402
from waldboost.model import Model as RModel  # noqa: E402
from waldboost.training import DTree as RDTree  # noqa: E402

from waldboost_b200 import synthetic as S  # noqa: E402


def ref_trees(trees):
    return [RDTree([tuple(f) for f in t.feature], t.threshold, t.left, t.right, t.prediction) for t in trees]


def ref_model(shape, opts, trees, thetas):
    M = RModel(shape, opts)
    for t, th in zip(ref_trees(trees), thetas):
        M.append(t, float(th))
    return M


def calibrate(M, chns_list, T, keep_total, subsample=None, seed=0, recipe=None):
    """wald thetas from the reference's own DTree.predict_on_image over all windows of the given maps
    (or a seeded random fraction `subsample` of them)."""
    m, n, _ = M.shape
    maps, R, Cc = [], [], []
    rng = np.random.default_rng(seed)
    for k, X in enumerate(chns_list):
        u, v, _ = X.shape
        rs, cs = np.indices((max(u - m, 0), max(v - n, 0)))
        rs, cs = rs.ravel(), cs.ravel()
        if subsample is not None:
            pick = rng.random(rs.size) < subsample
            rs, cs = rs[pick], cs[pick]
        maps.append(np.full(rs.size, k)); R.append(rs); Cc.append(cs)
    mp, R, Cc = np.concatenate(maps), np.concatenate(R), np.concatenate(Cc)

    def stage(t, alive):
        idx = np.arange(mp.size) if alive is None else alive
        out = np.empty(idx.size, np.float32)
        for k, X in enumerate(chns_list):
            sel = mp[idx] == k
            out[sel] = M.classifier[t].predict_on_image(X, R[idx][sel], Cc[idx][sel])
        return out
    return (recipe or S.calibrate_thetas)(stage, T, keep_total)


def detect_record(M, image):
    M.reset()
    levels = []
    for chns, scale in M.channels(image):
        r, c, h = M.predict_on_image(chns)
        levels.append((r, c, h, scale, chns))
    M2 = M
    n_loc, n_weak = M.n_loc, M.n_weak
    M.reset()
    dt = M2.detect(image)
    return levels, dt.get(), dt.get_field("scores"), n_loc, n_weak


def config_B():
    """BASELINE config B / E model: 12x12x4 grad_hist, 1024 depth-2 stages, 'wald' thetas calibrated with the
    reference on a seeded 8 % sample of the windows of all levels of the 1080p frames seed 1000..1003 (SURVEY.md 8d
    calibrates on frame 1000 only; four frames make the thresholds robust to the frames' different noise amplitudes) -> configB_model.pb (used by bench.py), plus the
    reference's detect() output on a 540x960 crop of frame 1001 (a 1080p reference run costs minutes)."""
    shape = (12, 12, 4)
    opts = dict(shrink=2, n_per_oct=8, smooth=1, channels=rch.grad_hist)
    lv = []
    for seed in range(1000, 1004):
        lv += [c for c, _ in rch.channel_pyramid(S.synthetic_frame(seed, 1080, 1920), opts) if c.shape[0] > 12 and c.shape[1] > 12]
    first = [c for c in lv if c.shape[:2] == (540, 960)]
    lo, hi = S.channel_quantiles(np.concatenate([c.reshape(-1, 4) for c in first])[None])
    trees = S.random_trees(shape, 1024, 2, lo, hi, seed=7)
    M = ref_model(shape, opts, trees, [-np.inf] * 1024)
    th = calibrate(M, lv, 1024, 1e-4, subsample=0.08, seed=1)
    M.theta = [float(x) for x in th]
    M.save(os.path.join(HERE, "configB_model.pb"))
    config_B_detect()


def config_B_detect():
    """the reference's detect() with the committed config B model on the centre 540x960 crop of frame 1001."""
    M = RModel.load(os.path.join(HERE, "configB_model.pb"))
    crop = np.ascontiguousarray(S.synthetic_frame(1001, 1080, 1920)[270:810, 480:1440])
    levels, boxes, scores, n_loc, n_weak = detect_record(M, crop)
    np.savez_compressed(os.path.join(HERE, "configB_detect.npz"), boxes=boxes, scores=scores, n_loc=np.int64(n_loc),
                        n_weak=np.int64(n_weak), level_counts=np.array([r.size for r, *_ in levels]))
    print("config B (540x960 crop): hits", scores.size, "n_loc", n_loc, "n_weak", n_weak, "eval_cost", n_weak / n_loc)


def fpga_pyramid():
    """integer channels of the reference's FPGA variant (waldboost/fpga/channels.py) through channel_pyramid."""
    from waldboost.fpga import channels as fch
    frame = S.synthetic_frame(1000, 96, 128)
    noise = S.noise_frame(3, 70, 90)
    cfgs = {"hist4u1_s2_sm1": dict(shrink=2, n_per_oct=4, smooth=1, channels=fch.grad_hist_4_u1),
            "hist4u1_s1_sm0": dict(shrink=1, n_per_oct=2, smooth=0, channels=fch.grad_hist_4_u1),
            "magu1_s2_sm1": dict(shrink=2, n_per_oct=3, smooth=1, channels=fch.grad_mag_u1),
            "magu1_s1_sm1": dict(shrink=1, n_per_oct=2, smooth=1, channels=fch.grad_mag_u1)}
    out = {"frame": frame, "noise": noise}
    for name, opts in cfgs.items():
        for tag, img in (("frame", frame), ("noise", noise)):
            for k, (chns, scale) in enumerate(rch.channel_pyramid(img, opts)):
                out[f"{name}/{tag}/{k}"] = chns
                out[f"{name}/{tag}/{k}/scale"] = np.float64(scale)
    np.savez_compressed(os.path.join(HERE, "fpga_pyramid.npz"), **out)
    print("fpga fixtures:", len(out))


def multi_model():
    """waldboost.detect(image, A, B, response_scale=[1.0, 0.5]) of the reference (__init__.py:75-130): one shared
    pyramid, two models of different window sizes, scores scaled per model, `label` = model index; the output order is
    (level, model, r, c).  ref_harness restores the `np.int` alias that __init__.py:128 still names."""
    frame = S.synthetic_frame(1000, 200, 260)
    opts = dict(shrink=2, n_per_oct=8, smooth=1, channels=rch.grad_hist)
    lv = list(rch.channel_pyramid(frame, opts))
    lo, hi = S.channel_quantiles(lv[0][0])
    models = []
    for name, shape, seed in (("A", (12, 12, 4), 1), ("B", (16, 10, 4), 2)):
        trees = S.random_trees(shape, 16, 2, lo, hi, seed=seed)
        M = ref_model(shape, opts, trees, [-np.inf] * 16)
        M.theta = [float(x) for x in calibrate(M, [lv[0][0]], 16, 1e-2)]
        M.save(os.path.join(HERE, f"multi_{name}_model.pb"))
        models.append(RModel.load(os.path.join(HERE, f"multi_{name}_model.pb")))
    dt = wb.detect(frame, *models, response_scale=[1.0, 0.5])
    one = [m.detect(frame) for m in models]
    np.savez_compressed(os.path.join(HERE, "multi_detect.npz"), boxes=dt.get(), scores=dt.get_field("scores"),
                        label=np.asarray(dt.get_field("label"), np.int64),
                        boxes_A=one[0].get(), scores_A=one[0].get_field("scores"),
                        boxes_B=one[1].get(), scores_B=one[1].get_field("scores"))
    print("multi-model: hits", len(dt), "labels", np.bincount(np.asarray(dt.get_field("label"), np.int64)))


def _mag_hist(im):
    """config C's 10-channel feature: the concatenation of the reference's own two functions on the same image."""
    return np.concatenate([rch.grad_mag(im), rch.grad_hist(im, 9)], axis=-1)


def config_C():
    """BASELINE config C at its stated size: one 3840x2160 uint8 frame, 20x20x10 model (grad_mag + 9-bin grad_hist,
    shrink 2, n_per_oct 8, smooth 1), 256 depth-2 stages with wald thetas calibrated by the reference on a seeded 2 %
    sample of the windows of every level.  Minutes of CPU."""
    import waldboost_b200 as wbp
    shape = (20, 20, 10)
    opts = dict(shrink=2, n_per_oct=8, smooth=1, channels=_mag_hist)
    frame = S.synthetic_frame(1000, 2160, 3840)
    lv = [c for c, _ in rch.channel_pyramid(frame, opts) if c.shape[0] > 20 and c.shape[1] > 20]
    lo, hi = S.channel_quantiles(lv[0])
    trees = S.random_trees(shape, 256, 2, lo, hi, seed=7)
    M = ref_model(shape, opts, trees, [-np.inf] * 256)
    th = calibrate(M, lv, 256, 1e-4, subsample=0.02, seed=1)
    M.theta = [float(x) for x in th]
    # the model file is written by the product package (the reference cannot name a lambda as channel function)
    P = wbp.Model(shape, dict(shrink=2, n_per_oct=8, smooth=1, channels=wbp.channels.grad_mag_hist))
    for t, thv in zip(trees, th):
        P.append(t, float(np.float32(thv)))
    P.save(os.path.join(HERE, "configC_model.pb"))
    P = wbp.Model.load(os.path.join(HERE, "configC_model.pb"))
    M = ref_model(shape, opts, P.classifier, P.theta)          # float32-exact thresholds and thetas on both sides
    levels, boxes, scores, n_loc, n_weak = detect_record(M, frame)
    np.savez_compressed(os.path.join(HERE, "configC_detect.npz"), boxes=boxes, scores=scores, n_loc=np.int64(n_loc),
                        n_weak=np.int64(n_weak), level_counts=np.array([r.size for r, *_ in levels]))
    print("config C (3840x2160): hits", scores.size, "n_loc", n_loc, "n_weak", n_weak, "eval_cost", n_weak / n_loc)


def config_D():
    """BASELINE config D's cascade shape at full length: 2048 depth-4 stages (12x12x4 grad_hist) through the reference's
    Model.scan_channels on one 200x260 frame -- per level the surviving (r, c, h)."""
    shape = (12, 12, 4)
    opts = dict(shrink=2, n_per_oct=8, smooth=1, channels=rch.grad_hist)
    frame = S.synthetic_frame(1003, 200, 260)
    lv = [c for c, _ in rch.channel_pyramid(frame, opts) if c.shape[0] > 12 and c.shape[1] > 12]
    lo, hi = S.channel_quantiles(lv[0])
    trees = S.random_trees(shape, 2048, 4, lo, hi, seed=7)
    M = ref_model(shape, opts, trees, [-np.inf] * 2048)
    th = calibrate(M, lv[:4], 2048, 1e-3)
    M.theta = [float(x) for x in th]
    M.save(os.path.join(HERE, "configD_model.pb"))
    M = RModel.load(os.path.join(HERE, "configD_model.pb"))
    M.reset()
    rec = {}
    k = 0
    for k, (chns, scale, (r, c, h)) in enumerate(M.scan_channels(frame)):
        rec[f"{k}/r"], rec[f"{k}/c"], rec[f"{k}/h"] = r.astype(np.int32), c.astype(np.int32), h
        rec[f"{k}/scale"] = np.float64(scale)
    rec["n_levels"] = np.int64(k + 1)
    rec["n_loc"], rec["n_weak"] = np.int64(M.n_loc), np.int64(M.n_weak)
    np.savez_compressed(os.path.join(HERE, "configD_scan.npz"), **rec)
    print("config D (2048 x depth 4): levels", k + 1, "survivors", sum(rec[f"{i}/r"].size for i in range(k + 1)),
          "n_loc", M.n_loc, "n_weak", M.n_weak, "eval_cost", M.n_weak / M.n_loc)


def main():
    if "--fpga" in sys.argv:
        return fpga_pyramid()
    if "--config-b" in sys.argv:
        return config_B()
    if "--config-b-detect" in sys.argv:
        return config_B_detect()
    if "--multi" in sys.argv:
        return multi_model()
    if "--config-c" in sys.argv:
        return config_C()
    if "--config-d" in sys.argv:
        return config_D()
    # ---------------------------------------------------------------- small pyramid fixtures
    frame = S.synthetic_frame(1000, 96, 128)
    frame_f = frame.astype(np.float32) + np.random.default_rng(5).random(frame.shape).astype(np.float32)
    cfgs = {
        "hist4_s2_sm1": (dict(shrink=2, n_per_oct=4, smooth=1, channels=rch.grad_hist)),
        "hist4_s1_sm0": (dict(shrink=1, n_per_oct=2, smooth=0, channels=rch.grad_hist)),
        "mag_s2_sm1": (dict(shrink=2, n_per_oct=3, smooth=1, channels=rch.grad_mag)),
        "hist6full_s2_sm1": (dict(shrink=2, n_per_oct=2, smooth=1, channels=lambda im: rch.grad_hist(im, 6, True, 2))),
    }
    out = {"frame_u8": frame, "frame_f32": frame_f}
    for name, opts in cfgs.items():
        for tag, img in (("u8", frame), ("f32", frame_f)):
            for k, (chns, scale) in enumerate(rch.channel_pyramid(img, opts)):
                out[f"{name}/{tag}/{k}"] = chns
                out[f"{name}/{tag}/{k}/scale"] = np.float64(scale)
    np.savez_compressed(os.path.join(HERE, "small_pyramid.npz"), **out)

    # ---------------------------------------------------------------- small cascade fixtures
    opts = dict(shrink=2, n_per_oct=4, smooth=1, channels=rch.grad_hist)
    shape = (12, 12, 4)
    lv = list(rch.channel_pyramid(frame, opts))
    lo, hi = S.channel_quantiles(lv[0][0])
    trees = S.random_trees(shape, 24, 2, lo, hi, seed=7)
    M = ref_model(shape, opts, trees, [-np.inf] * 24)
    th = calibrate(M, [lv[0][0], lv[1][0]], 24, 1e-2, recipe=S.calibrate_thetas_v0)
    M.theta = [float(x) for x in th]
    M.save(os.path.join(HERE, "small_model.pb"))
    M = RModel.load(os.path.join(HERE, "small_model.pb"))
    rec = {}
    for prof in ("wald", "dense"):
        if prof == "dense":
            M.theta = [-np.inf] * 24
        levels, boxes, scores, n_loc, n_weak = detect_record(M, frame)
        rec[f"{prof}/boxes"], rec[f"{prof}/scores"] = boxes, scores
        rec[f"{prof}/n_loc"], rec[f"{prof}/n_weak"] = np.int64(n_loc), np.int64(n_weak)
        for k, (r, c, h, scale, _) in enumerate(levels):
            rec[f"{prof}/{k}/r"], rec[f"{prof}/{k}/c"], rec[f"{prof}/{k}/h"] = r.astype(np.int32), c.astype(np.int32), h
    np.savez_compressed(os.path.join(HERE, "small_detect.npz"), **rec)

    # ---------------------------------------------------------------- generic topologies
    rng = np.random.default_rng(11)
    gtrees = S.random_trees(shape, 6, 3, lo, hi, seed=3) + S.random_trees(shape, 6, 1, lo, hi, seed=4)
    # unbalanced: root -> (leaf, internal -> (leaf, leaf)) in pre-order
    for _ in range(6):
        feature = np.zeros((5, 3), np.uint8); threshold = np.full(5, -2, np.float32); pred = np.zeros(5, np.float32)
        for k in (0, 2):
            ch = int(rng.integers(0, 4))
            feature[k] = (rng.integers(0, 12), rng.integers(0, 12), ch)
            threshold[k] = np.float32(rng.uniform(lo[ch], hi[ch]))
        pred[[1, 3, 4]] = rng.normal(0, 0.5, 3).astype(np.float32)
        from waldboost_b200.training import DTree
        gtrees.append(DTree(feature, threshold, [1, -1, 3, -1, -1], [2, -1, 4, -1, -1], pred))
    order = rng.permutation(len(gtrees))
    gtrees = [gtrees[i] for i in order]
    G = ref_model(shape, opts, gtrees, [-np.inf] * len(gtrees))
    gth = calibrate(G, [lv[0][0]], len(gtrees), 3e-2, recipe=S.calibrate_thetas_v0)
    gth[::4] = -np.inf
    G.theta = [float(x) for x in gth]
    G.save(os.path.join(HERE, "generic_model.pb"))
    G = RModel.load(os.path.join(HERE, "generic_model.pb"))
    levels, boxes, scores, n_loc, n_weak = detect_record(G, frame)
    np.savez_compressed(os.path.join(HERE, "generic_detect.npz"), boxes=boxes, scores=scores, n_loc=np.int64(n_loc), n_weak=np.int64(n_weak))

    # ---------------------------------------------------------------- BASELINE config A
    optsA = dict(shrink=2, n_per_oct=8, smooth=1, channels=rch.grad_hist)
    frameA = S.synthetic_frame(1000, 480, 640)
    lvA = list(rch.channel_pyramid(frameA, optsA))
    loA, hiA = S.channel_quantiles(lvA[0][0])
    treesA = S.random_trees(shape, 256, 2, loA, hiA, seed=7)
    MA = ref_model(shape, optsA, treesA, [-np.inf] * 256)
    thA = calibrate(MA, [lvA[0][0]], 256, 1e-4, recipe=S.calibrate_thetas_v0)
    MA.theta = [float(x) for x in thA]
    MA.save(os.path.join(HERE, "configA_model.pb"))
    MA = RModel.load(os.path.join(HERE, "configA_model.pb"))
    levels, boxes, scores, n_loc, n_weak = detect_record(MA, frameA)
    np.savez_compressed(os.path.join(HERE, "configA_detect.npz"), boxes=boxes, scores=scores, n_loc=np.int64(n_loc),
                        n_weak=np.int64(n_weak), level_sums=np.array([c.astype(np.float64).sum() for *_, c in levels]),
                        level_counts=np.array([r.size for r, *_ in levels]), thr_lo=loA, thr_hi=hiA)
    print("config A: hits", scores.size, "n_loc", n_loc, "n_weak", n_weak, "eval_cost", n_weak / n_loc)


if __name__ == "__main__":
    main()
