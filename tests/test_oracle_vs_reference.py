"""CPU, build container only: the oracle (oracle/wb_oracle.py) against the UNMODIFIED reference imported live from
/root/reference under the shims of oracle/ref_harness.py, on seeded random inputs that are not among the committed
fixtures.  Skipped where the reference is absent (the GPU box)."""
import os

import numpy as np
import pytest

import ref_harness
import wb_oracle as O

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="/root/reference is not present")


@pytest.fixture(scope="module")
def ref():
    wb = ref_harness.import_reference()
    from waldboost import channels as rch
    from waldboost.model import Model as RModel
    from waldboost.training import DTree as RDTree
    return wb, rch, RModel, RDTree


def _frame(seed, h, w, dtype):
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w))
    for _ in range(4):
        y, x, s = rng.integers(0, h - 8), rng.integers(0, w - 8), rng.integers(6, 30)
        img[y:y + s, x:x + s] += rng.uniform(0.2, 1.0)
    img = np.clip(img + rng.random((h, w)) * 0.3, 0, 1) * 255
    out = img.astype(np.uint8)
    if dtype == np.float32:
        out = out.astype(np.float32) + rng.random((h, w)).astype(np.float32)
    return out


def _trees(rng, shape, T, depth, lo, hi):
    """full binary trees in sklearn pre-order layout."""
    m, n, C = shape
    out = []
    for _ in range(T):
        N = 2 ** (depth + 1) - 1
        feature = np.zeros((N, 3), np.uint8); thr = np.full(N, -2, np.float32); left = np.full(N, -1, np.int8)
        right = np.full(N, -1, np.int8); pred = np.zeros(N, np.float32)
        nxt = [0]

        def build(d):
            k = nxt[0]; nxt[0] += 1
            if d == depth:
                pred[k] = np.float32(rng.normal(0, 0.5))
                return k
            ch = int(rng.integers(0, C))
            feature[k] = (rng.integers(0, m), rng.integers(0, n), ch)
            thr[k] = np.float32(rng.uniform(lo[ch], hi[ch]))
            left[k] = build(d + 1)
            right[k] = build(d + 1)
            return k
        build(0)
        out.append((feature, thr, left, right, pred))
    return out


CONFIGS = [
    ("hist4", dict(shrink=2, n_per_oct=3, smooth=1), lambda rch: rch.grad_hist, O.grad_hist, 4),
    ("hist7_s1", dict(shrink=1, n_per_oct=2, smooth=0), lambda rch: (lambda im: rch.grad_hist(im, 7, False, 1)),
     lambda im: O.grad_hist(im, 7, False, 1), 7),
    ("mag3", dict(shrink=2, n_per_oct=2, smooth=1), lambda rch: (lambda im: rch.grad_mag(im, 3)), lambda im: O.grad_mag(im, 3), 1),
]


@pytest.mark.parametrize("dtype", [np.uint8, np.float32])
@pytest.mark.parametrize("cfg", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_pyramid_and_detect_match_live_reference(ref, cfg, dtype):
    wb, rch, RModel, RDTree = ref
    name, base, rfn, ofn, C = cfg
    seed = abs(hash((name, np.dtype(dtype).name))) % 10000
    rng = np.random.default_rng(seed)
    frame = _frame(seed, int(rng.integers(70, 130)), int(rng.integers(90, 170)), dtype)
    ropts, oopts = dict(base, channels=rfn(rch)), dict(base, channels=ofn)
    rl, ol = list(rch.channel_pyramid(frame, ropts)), list(O.channel_pyramid(frame, oopts))
    assert len(rl) == len(ol) > 0
    for (a, sa), (b, sb) in zip(rl, ol):
        assert sa == sb and a.dtype == b.dtype and np.array_equal(a, b)
    # a random cascade with a few finite thetas, depth 1..3
    X0 = ol[0][0]
    lo, hi = np.quantile(X0.reshape(-1, C), 0.1, axis=0), np.quantile(X0.reshape(-1, C), 0.9, axis=0)
    shape = (int(rng.integers(4, 13)), int(rng.integers(4, 13)), C)
    depth = int(rng.integers(1, 4))
    trees = _trees(rng, shape, 10, depth, lo, hi)
    thetas = [(-np.inf if k % 3 == 0 else float(np.float32(-0.3 * (k + 1) ** 0.5))) for k in range(10)]
    RM, OC = RModel(shape, ropts), O.Cascade(shape, oopts)
    for (f, t, l, r, p), th in zip(trees, thetas):
        RM.append(RDTree([tuple(x) for x in f], t, l, r, p), th)
        OC.append(O.DTree([tuple(x) for x in f], t, l, r, p), th)
    RM.reset()
    dt = RM.detect(frame)
    boxes, scores, _ = OC.detect(frame)
    assert np.array_equal(dt.get(), boxes) and np.array_equal(dt.get_field("scores"), scores)
    assert (RM.n_loc, RM.n_weak) == (OC.n_loc, OC.n_weak) and RM.n_loc > 0
    # per level: survivors and leaf values of one stage
    for (chns, scale, (r, c, h)), (ochns, oscale, (orr, oc, oh)) in zip(RM.scan_channels(frame), OC.scan_channels(frame)):
        assert np.array_equal(r, orr) and np.array_equal(c, oc) and np.array_equal(h, oh)


def test_multi_model_detect_matches_live_reference(ref):
    wb, rch, RModel, RDTree = ref
    rng = np.random.default_rng(77)
    frame = _frame(77, 120, 150, np.uint8)
    base = dict(shrink=2, n_per_oct=4, smooth=1)
    ropts, oopts = dict(base, channels=rch.grad_hist), dict(base, channels=O.grad_hist)
    X0 = next(iter(O.channel_pyramid(frame, oopts)))[0]
    lo, hi = np.quantile(X0.reshape(-1, 4), 0.1, axis=0), np.quantile(X0.reshape(-1, 4), 0.9, axis=0)
    RMs, OCs = [], []
    for shape in ((8, 8, 4), (12, 6, 4), (5, 14, 4)):
        trees = _trees(rng, shape, 6, 2, lo, hi)
        RM, OC = RModel(shape, ropts), O.Cascade(shape, oopts)
        for k, (f, t, l, r, p) in enumerate(trees):
            th = float(np.float32(-0.2 * (k + 1)))
            RM.append(RDTree([tuple(x) for x in f], t, l, r, p), th)
            OC.append(O.DTree([tuple(x) for x in f], t, l, r, p), th)
        RMs.append(RM); OCs.append(OC)
    dt = wb.detect(frame, *RMs, response_scale=[1.0, 0.5, 2.0])
    boxes, scores, label = O.detect_multi(frame, OCs, response_scale=[1.0, 0.5, 2.0])
    assert len(dt) > 0 and np.array_equal(dt.get(), boxes)
    assert np.array_equal(dt.get_field("scores"), scores) and np.array_equal(np.asarray(dt.get_field("label"), np.int64), label)
    with pytest.raises(ValueError):
        O.detect_multi(frame, OCs, response_scale=[1.0])


def test_make_golden_reproduces_committed_fixtures(ref, tmp_path):
    """the committed recipe regenerates the small / generic / config A fixtures (models byte for byte, arrays equal)."""
    import importlib.util
    import shutil
    import subprocess
    import sys
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    work = tmp_path / "golden"
    shutil.copytree(here, work)
    for f in ("small_model.pb", "small_detect.npz", "small_pyramid.npz", "generic_model.pb", "generic_detect.npz",
              "configA_model.pb", "configA_detect.npz"):
        os.remove(work / f)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.dirname(os.path.dirname(here)), os.environ.get("PYTHONPATH", "")]))
    # make_golden.py writes next to itself: run the copy
    root = os.path.dirname(os.path.dirname(here))
    code = ("import sys, runpy; sys.path.insert(0, %r); sys.path.insert(0, %r); sys.argv=['make_golden.py']; "
            "import types; g = runpy.run_path(%r, run_name='not_main'); "
            "g['HERE'] = %r; g['main'].__globals__['HERE'] = %r; g['main']()" %
            (root, os.path.join(root, "oracle"), str(work / "make_golden.py"), str(work), str(work)))
    subprocess.run([sys.executable, "-c", code], check=True, env=env, cwd=root, timeout=1500)
    for f in ("small_model.pb", "generic_model.pb", "configA_model.pb"):
        assert open(work / f, "rb").read() == open(os.path.join(here, f), "rb").read(), f
    for f in ("small_detect.npz", "small_pyramid.npz", "generic_detect.npz", "configA_detect.npz"):
        a, b = np.load(work / f), np.load(os.path.join(here, f))
        assert set(a.files) == set(b.files), f
        for k in a.files:
            assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k]), (f, k)
