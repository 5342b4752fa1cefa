"""CPU: pin the oracle (oracle/wb_oracle.py) against (1) the dependencies whose arithmetic the reference
delegates to (scipy.ndimage, numba, numpy -- all in the image), (2) hand-computable known answers from SURVEY.md
sections 0 and 8c, and (3) golden outputs of the unmodified reference (tests/golden/, made by make_golden.py)."""
import os
import warnings

import numpy as np
import pytest

import wb_oracle as O
from helpers import GOLDEN

rng = np.random.default_rng(123)


# ------------------------------------------------------------------------------------- dependency pins
@pytest.mark.parametrize("shape", [(33, 47), (8, 8), (5, 64), (1, 9), (3, 3)])
def test_correlate_matches_scipy(shape):
    from scipy.ndimage import convolve1d
    x = (rng.random(shape) * 255).astype(np.float32)
    H = np.array([1, 2, 1], "f4")
    D = np.array([-1, 0, 1], "f4")
    gy = convolve1d(convolve1d(x, H, axis=1), D, axis=0)
    gx = convolve1d(convolve1d(x, H, axis=0), D, axis=1)
    ogx, ogy = O.gradients(x)
    assert np.array_equal(gx, ogx) and np.array_equal(gy, ogy)
    tri = O.triangle_kernel(5)
    out = convolve1d(x, tri, axis=0)
    convolve1d(out, tri, axis=1, output=out)
    assert np.array_equal(out, O.separable_convolve(x, tri))


@pytest.mark.parametrize("dtype", [np.uint8, np.float32])
@pytest.mark.parametrize("size", [((135, 240), (122, 218)), ((480, 640), (440, 586)), ((67, 120), (66, 120)),
                                  ((33, 60), (18, 32)), ((16, 30), (16, 30)), ((1080, 1920), (990, 1760))])
def test_resize_matches_scipy_zoom(dtype, size):
    import scipy.ndimage as ndi
    (h, w), (nh, nw) = size
    img = (rng.random((h, w)) * 255).astype(np.uint8).astype(dtype)
    if dtype == np.float32:
        img += rng.random((h, w)).astype(np.float32)
    src = img if img.dtype.char in "df" else img.astype(np.float64)
    factors = np.divide(src.shape, (nh, nw))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = ndi.zoom(src, [1 / f for f in factors], order=1, mode="mirror", cval=0, grid_mode=True)
    ref = np.clip(ref, src.min(), src.max()).astype(dtype)
    assert np.array_equal(O.resize_bilinear(img, nh, nw), ref)


def test_resize_flat_and_saturated_regions():
    """truncation cliff: flat regions make the float64 sum land a hair under the integer."""
    import scipy.ndimage as ndi
    img = np.full((135, 240), 200, np.uint8)
    img[:40] = 255
    img[100:] = 7
    src = img.astype(np.float64)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = ndi.zoom(src, [1 / f for f in np.divide(src.shape, (122, 218))], order=1, mode="mirror", grid_mode=True)
    ref = np.clip(ref, src.min(), src.max()).astype(np.uint8)
    assert np.array_equal(O.resize_bilinear(img, 122, 218), ref)


def test_numba_semantics():
    import numba as nb

    @nb.njit
    def pool(arr):
        u, v = arr.shape[0], arr.shape[1]
        ul, vl = u - (u % 2), v - (v % 2)
        return ((arr[0:ul:2, 0:vl:2, ...] + arr[1:ul:2, 0:vl:2, ...] + arr[0:ul:2, 1:vl:2, ...] + arr[1:ul:2, 1:vl:2, ...]) / 4).astype(arr.dtype)

    @nb.stencil(neighborhood=((-1, 1), (-1, 1), (0, 0)))
    def sm(arr):
        return arr[-1, -1, 0] + 2 * arr[-1, 0, 0] + arr[-1, 1, 0] + 2 * arr[0, -1, 0] + 4 * arr[0, 0, 0] + 2 * arr[0, 1, 0] + \
            arr[1, -1, 0] + 2 * arr[1, 0, 0] + arr[1, 1, 0]

    @nb.njit
    def smooth(arr):
        out = np.empty_like(arr)
        out[:] = sm(arr) / 16
        return out

    u8 = (rng.random((33, 47)) * 255).astype(np.uint8)
    f3 = (rng.random((21, 18, 4)) * 100).astype(np.float32)
    assert np.array_equal(pool(u8), O.avg_pool_2(u8))
    assert np.array_equal(pool(f3), O.avg_pool_2(f3))
    assert np.array_equal(smooth(f3), O.smooth_image_3d(f3))
    tiny = np.ones((2, 4, 3), np.float32)
    assert np.array_equal(smooth(tiny), O.smooth_image_3d(tiny)) and not O.smooth_image_3d(tiny).any()


# ------------------------------------------------------------------------------------- known answers
def test_kat_gradients_ramp():
    ramp = (np.arange(25, dtype=np.float32).reshape(5, 5))
    gx, gy = O.gradients(ramp)
    assert gx[2, 2] == -8 and gx[2, 0] == -4 and gx[2, 4] == -4
    assert gy[2, 2] == -40 and gy[0, 2] == -20 and gy[4, 2] == -20


def test_kat_pool_and_smooth():
    assert O.avg_pool_2(np.array([[10, 20], [30, 41]], np.uint8))[0, 0] == 25
    assert O.avg_pool_2(np.array([[10, 20], [30, 41]], np.float32))[0, 0] == 25.25
    assert O.avg_pool_2(np.array([[250, 250], [250, 250]], np.uint8))[0, 0] == 250      # no uint8 wrap-around
    assert O.avg_pool_2(np.zeros((7, 9, 3), np.float32)).shape == (3, 4, 3)
    s = O.smooth_image_3d(np.ones((5, 6, 2), np.float32))
    assert (s[1:-1, 1:-1] == 1).all() and s[0].sum() == 0 and s[:, 0].sum() == 0 and s[-1].sum() == 0 and s[:, -1].sum() == 0
    assert np.allclose(O.triangle_kernel(5) * 36, [1, 2, 3, 4, 5, 6, 5, 4, 3, 2, 1])


def test_kat_window_grid_and_rejection():
    """facts 1, 7, 8: grid (u-m)x(v-n); left iff x <= thr; keep iff hs >= theta; -inf = no test."""
    X = np.zeros((6, 7, 1), np.float32)
    X[2, 3, 0] = 5.0
    tree = O.DTree([(2, 3, 0), None, None], [4.0, -2, -2], [1, -1, -1], [2, -1, -1], [0, -1.0, 2.0])
    Cs = O.Cascade((4, 4, 1), None)
    Cs.append(tree, -np.inf)
    r, c, h = Cs.predict_on_image(X)
    assert r.size == (6 - 4) * (7 - 4) and Cs.n_loc == 6 and Cs.n_weak == 6
    assert h[(r == 0) & (c == 0)][0] == 2.0 and (h == 2.0).sum() == 1   # only window (0,0) sees X[2,3]
    Cs.theta = [2.0]
    r, c, h = Cs.predict_on_image(X)
    assert list(zip(r, c)) == [(0, 0)]
    Xeq = np.full((6, 7, 1), 4.0, np.float32)     # x == thr goes left
    Cs.theta = [-np.inf]
    assert (Cs.predict_on_image(Xeq)[2] == -1.0).all()


def test_level_sizes_match_survey():
    sizes = [O.level_size(1080, 1920, i, 8, 2) for i in range(8)]
    assert [(a // 2, b // 2) for a, b in sizes] == [(540, 960), (495, 880), (454, 807), (416, 740), (381, 678),
                                                    (350, 622), (321, 570), (294, 523)]
    octs = [o.shape for o in O.image_octaves(np.zeros((1080, 1920), np.uint8))]
    assert octs == [(1080, 1920), (540, 960), (270, 480), (135, 240), (67, 120), (33, 60), (16, 30), (8, 15)]


# ------------------------------------------------------------------------------------- golden vectors of the reference
CFG = {
    "hist4_s2_sm1": dict(shrink=2, n_per_oct=4, smooth=1, channels=O.grad_hist),
    "hist4_s1_sm0": dict(shrink=1, n_per_oct=2, smooth=0, channels=O.grad_hist),
    "mag_s2_sm1": dict(shrink=2, n_per_oct=3, smooth=1, channels=O.grad_mag),
    "hist6full_s2_sm1": dict(shrink=2, n_per_oct=2, smooth=1, channels=lambda im: O.grad_hist(im, 6, True, 2)),
}


@pytest.mark.parametrize("name", list(CFG))
@pytest.mark.parametrize("tag", ["u8", "f32"])
def test_oracle_pyramid_equals_reference_golden(name, tag):
    g = np.load(os.path.join(GOLDEN, "small_pyramid.npz"))
    img = g["frame_" + tag]
    levels = list(O.channel_pyramid(img, CFG[name]))
    n_ref = len([k for k in g.files if k.startswith(f"{name}/{tag}/") and k.endswith("/scale")])
    assert len(levels) == n_ref > 0
    for k, (chns, scale) in enumerate(levels):
        ref = g[f"{name}/{tag}/{k}"]
        assert chns.dtype == np.float32 and chns.shape == ref.shape
        assert np.array_equal(chns, ref), f"level {k}"
        assert scale == float(g[f"{name}/{tag}/{k}/scale"])


def _load_oracle_model(path):
    import waldboost_b200 as wb
    from helpers import oracle_cascade
    return oracle_cascade(wb.Model.load(path))


@pytest.mark.parametrize("prof", ["wald", "dense"])
def test_oracle_cascade_equals_reference_golden(prof):
    g = np.load(os.path.join(GOLDEN, "small_detect.npz"))
    frame = np.load(os.path.join(GOLDEN, "small_pyramid.npz"))["frame_u8"]
    Cs = _load_oracle_model(os.path.join(GOLDEN, "small_model.pb"))
    if prof == "dense":
        Cs.theta = [-np.inf] * len(Cs)
    for k, (chns, scale) in enumerate(Cs.channels(frame)):
        r, c, h = Cs.predict_on_image(chns)
        assert np.array_equal(r, g[f"{prof}/{k}/r"]) and np.array_equal(c, g[f"{prof}/{k}/c"])
        assert np.array_equal(h, g[f"{prof}/{k}/h"])
    assert Cs.n_loc == int(g[f"{prof}/n_loc"]) and Cs.n_weak == int(g[f"{prof}/n_weak"])
    Cs.reset()
    boxes, scores, _ = Cs.detect(frame)
    assert np.array_equal(boxes, g[f"{prof}/boxes"]) and np.array_equal(scores, g[f"{prof}/scores"])


def test_oracle_generic_topology_equals_reference_golden():
    g = np.load(os.path.join(GOLDEN, "generic_detect.npz"))
    frame = np.load(os.path.join(GOLDEN, "small_pyramid.npz"))["frame_u8"]
    Cs = _load_oracle_model(os.path.join(GOLDEN, "generic_model.pb"))
    boxes, scores, _ = Cs.detect(frame)
    assert np.array_equal(boxes, g["boxes"]) and np.array_equal(scores, g["scores"])
    assert Cs.n_loc == int(g["n_loc"]) and Cs.n_weak == int(g["n_weak"])


def test_oracle_config_A_equals_reference_golden():
    """BASELINE config A: 640x480, 12x12x4 grad_hist, 256 depth-2 stages -- full detect() of the reference."""
    from waldboost_b200 import synthetic as S
    g = np.load(os.path.join(GOLDEN, "configA_detect.npz"))
    Cs = _load_oracle_model(os.path.join(GOLDEN, "configA_model.pb"))
    frame = S.synthetic_frame(1000, 480, 640)
    sums, counts = [], []
    B, Sc = [], []
    for chns, scale in Cs.channels(frame):
        r, c, h = Cs.predict_on_image(chns)
        sums.append(chns.astype(np.float64).sum())
        counts.append(r.size)
        B.append(Cs.get_boxes(r, c, scale)); Sc.append(h)
    assert np.array_equal(np.array(sums), g["level_sums"]) and np.array_equal(np.array(counts), g["level_counts"])
    assert np.array_equal(np.concatenate(B), g["boxes"]) and np.array_equal(np.concatenate(Sc), g["scores"])
    assert Cs.n_loc == int(g["n_loc"]) == 407350 and Cs.n_weak == int(g["n_weak"])


def test_oracle_config_B_model_equals_reference_golden():
    """BASELINE config B model (1024 stages) on a 540x960 crop: oracle == reference detect()."""
    from waldboost_b200 import synthetic as S
    g = np.load(os.path.join(GOLDEN, "configB_detect.npz"))
    Cs = _load_oracle_model(os.path.join(GOLDEN, "configB_model.pb"))
    crop = np.ascontiguousarray(S.synthetic_frame(1001, 1080, 1920)[270:810, 480:1440])
    boxes, scores, _ = Cs.detect(crop)
    assert np.array_equal(boxes, g["boxes"]) and np.array_equal(scores, g["scores"])
    assert (Cs.n_loc, Cs.n_weak) == (int(g["n_loc"]), int(g["n_weak"]))


FPGA_CFG = {
    "hist4u1_s2_sm1": dict(shrink=2, n_per_oct=4, smooth=1, channels=O.grad_hist_4_u1),
    "hist4u1_s1_sm0": dict(shrink=1, n_per_oct=2, smooth=0, channels=O.grad_hist_4_u1),
    "magu1_s2_sm1": dict(shrink=2, n_per_oct=3, smooth=1, channels=O.grad_mag_u1),
    "magu1_s1_sm1": dict(shrink=1, n_per_oct=2, smooth=1, channels=O.grad_mag_u1),
}


@pytest.mark.parametrize("name", list(FPGA_CFG))
@pytest.mark.parametrize("tag", ["frame", "noise"])
def test_oracle_fpga_channels_equal_reference_golden(name, tag):
    """integer channels of the reference's FPGA variant (waldboost/fpga/channels.py) through channel_pyramid."""
    g = np.load(os.path.join(GOLDEN, "fpga_pyramid.npz"))
    levels = list(O.channel_pyramid(g[tag], FPGA_CFG[name]))
    n_ref = len([k for k in g.files if k.startswith(f"{name}/{tag}/") and k.endswith("/scale")])
    assert len(levels) == n_ref > 0
    for k, (chns, scale) in enumerate(levels):
        ref = g[f"{name}/{tag}/{k}"]
        assert chns.dtype == np.uint8 == ref.dtype and np.array_equal(chns, ref), f"level {k}"
        assert scale == float(g[f"{name}/{tag}/{k}/scale"])


# ------------------------------------------------------------------------------------- goldens added in round 2
def _load_cascade(path):
    """reference-written .pb -> oracle Cascade (host-side parsing by the product package's wire-format mirror)."""
    import waldboost_b200 as wb
    return O.cascade_from_model(wb.Model.load(path))


def test_multi_model_detect_vs_reference_golden():
    """waldboost.detect(image, A, B, response_scale=[1, .5]) of the reference: boxes, scores and labels in its order."""
    from waldboost_b200 import synthetic as S
    g = np.load(os.path.join(GOLDEN, "multi_detect.npz"))
    A, B = _load_cascade(os.path.join(GOLDEN, "multi_A_model.pb")), _load_cascade(os.path.join(GOLDEN, "multi_B_model.pb"))
    frame = S.synthetic_frame(1000, 200, 260)
    boxes, scores, label = O.detect_multi(frame, [A, B], response_scale=[1.0, 0.5])
    assert np.array_equal(boxes, g["boxes"]) and np.array_equal(scores, g["scores"]) and np.array_equal(label, g["label"])
    assert (label == 0).sum() > 0 and (label == 1).sum() > 0
    ba, sa, _ = A.detect(frame)
    assert np.array_equal(ba, g["boxes_A"]) and np.array_equal(sa, g["scores_A"])


def test_config_D_2048_depth4_scan_vs_reference_golden():
    """2048 depth-4 stages through scan_channels (BASELINE config D's cascade at full length)."""
    from waldboost_b200 import synthetic as S
    g = np.load(os.path.join(GOLDEN, "configD_scan.npz"))
    Cs = _load_cascade(os.path.join(GOLDEN, "configD_model.pb"))
    assert len(Cs) == 2048 and all(len(w.left) == 31 for w in Cs.classifier)
    frame = S.synthetic_frame(1003, 200, 260)
    n = 0
    for k, (chns, scale, (r, c, h)) in enumerate(Cs.scan_channels(frame)):
        assert scale == float(g[f"{k}/scale"])
        assert np.array_equal(r, g[f"{k}/r"]) and np.array_equal(c, g[f"{k}/c"]) and np.array_equal(h, g[f"{k}/h"])
        n += r.size
    assert k + 1 == int(g["n_levels"]) and n > 0
    assert (Cs.n_loc, Cs.n_weak) == (int(g["n_loc"]), int(g["n_weak"]))
