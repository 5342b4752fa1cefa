"""CPU: the C-ABI library loads and exports every symbol include/wbg.h declares; host-side geometry, model I/O and
error behaviour.  No compute call needs a GPU here."""
import ctypes as C
import functools
import os
import re
import zlib

import numpy as np
import pytest

import wb_oracle as O
import waldboost_b200 as wb
from waldboost_b200 import _native as N
from waldboost_b200 import synthetic as S
from waldboost_b200.engine import make_channel_opts, plan_geometry
from helpers import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPTS = dict(shrink=2, n_per_oct=8, smooth=1, channels=wb.channels.grad_hist)


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "wbg.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(wbg_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(N.SYMBOLS), declared ^ set(N.SYMBOLS)
    L = N.lib()
    for name in declared:
        assert getattr(L, name) is not None
    assert L.wbg_abi_version() == N.ABI_VERSION
    assert C.sizeof(N.Level) == 64 and C.sizeof(N.PlanInfo) == 64 and N.HIT_DTYPE.itemsize == 36


@pytest.mark.parametrize("size,win,C_,npo,shrink,n_loc", [
    ((480, 640), (12, 12), 4, 8, 2, 407350), ((1080, 1920), (12, 12), 4, 8, 2, 3045278),
    ((2160, 3840), (20, 20), 10, 8, 2, 12324184), ((96, 128), (12, 12), 4, 4, 2, None), ((333, 517), (9, 15), 4, 5, 1, None)])
def test_plan_geometry_matches_reference_arithmetic(size, win, C_, npo, shrink, n_loc):
    """level sizes / scales are the Python-double arithmetic of channels.py:124-131; window grid of model.py:243."""
    Hh, Ww = size
    fn = wb.channels.grad_hist if C_ == 4 else wb.channels.grad_mag_hist
    opts = dict(shrink=shrink, n_per_oct=npo, smooth=1, channels=fn)
    plan = plan_geometry(Hh, Ww, opts, wb.channels.resolve_channels(fn), *win)
    assert plan.C == C_
    octs = [o.shape for o in O.image_octaves(np.zeros(size, np.uint8))]
    assert plan.info.n_octaves == len(octs) and plan.n_levels == len(octs) * npo
    k, total, off = 0, 0, 0
    for (h, w) in octs:
        for i in range(npo):
            nh, nw = O.level_size(h, w, i, npo, shrink)
            lv = plan.levels[k]
            assert (lv.src_h, lv.src_w, lv.nh, lv.nw, lv.u, lv.v) == (h, w, nh, nw, nh // shrink, nw // shrink)
            assert lv.scale == nw / Ww / shrink
            assert (lv.win_rows, lv.win_cols) == (max(lv.u - win[0], 0), max(lv.v - win[1], 0))
            assert lv.chn_off == off and lv.win_off % 32 == 0
            off += -(-lv.u * lv.v * C_ // 4) * 4
            total += lv.win_rows * lv.win_cols
            k += 1
    assert plan.info.n_loc == total and plan.chn_floats == off
    if n_loc is not None:
        assert total == n_loc


def test_survey_byte_model():
    """channel bytes per frame that bench.py's roofline uses."""
    spec = wb.channels.resolve_channels(wb.channels.grad_hist)
    for size, expect in (((480, 640), 7_678_224), ((1080, 1920), 52_029_776)):
        plan = plan_geometry(*size, OPTS, spec, 12, 12)
        assert 4 * plan.C * sum(lv.u * lv.v for lv in plan.levels) == expect


def test_plan_rejects_bad_options():
    spec = wb.channels.resolve_channels(wb.channels.grad_hist)
    with pytest.raises(AssertionError):
        plan_geometry(64, 64, dict(OPTS, shrink=3), spec)
    with pytest.raises(N.WbgError):
        plan_geometry(64, 64, dict(OPTS, n_per_oct=0), spec)
    assert plan_geometry(7, 64, OPTS, spec).n_levels == 0


def test_no_gpu_means_loud_failure():
    """there is no CPU fallback: without a device the engine, plans with device tables and models all fail."""
    L = N.lib()
    if L.wbg_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from waldboost_b200.engine import ModelHandle, Plan, get_engine
    with pytest.raises(RuntimeError):
        get_engine()
    spec = wb.channels.resolve_channels(wb.channels.grad_hist)
    with pytest.raises(N.WbgError) as e:
        Plan(64, 64, make_channel_opts(OPTS, spec), 12, 12, device_tables=True)
    assert e.value.code == N.WBG_ECUDA
    M = wb.Model.load(os.path.join(GOLDEN, "small_model.pb"))
    with pytest.raises(N.WbgError):
        ModelHandle(M.shape, M.classifier, M.theta)
    with pytest.raises(RuntimeError):
        M.detect(np.zeros((64, 64), np.uint8))


def test_channel_function_resolution():
    r = wb.channels.resolve_channels
    assert r(wb.channels.grad_hist)["n_bins"] == 4 and r(wb.channels.grad_mag)["norm"] == 5
    assert r(functools.partial(wb.channels.grad_hist, n_bins=6, full=True))["full"] is True
    for bad in (lambda im: im, np.mean, functools.partial(wb.channels.grad_hist, 3), functools.partial(wb.channels.grad_hist, norm=2)):
        with pytest.raises(TypeError):
            r(bad)
    o = make_channel_opts(OPTS, r(wb.channels.grad_hist))
    theta = np.linspace(0, np.pi, 5)[:-1]
    assert [o.cos_t[i] for i in range(4)] == list(np.cos(theta)) and o.cos_t[2] == 6.123233995736766e-17


def test_pb_roundtrip_and_reference_compatibility(tmp_path):
    """files written by the reference load here; what we write is the same message (field for field)."""
    from waldboost_b200 import model_pb2
    src = os.path.join(GOLDEN, "small_model.pb")
    M = wb.load(src)
    assert M.shape == (12, 12, 4) and len(M) == 24 and M.channel_opts["channels"] is wb.channels.grad_hist
    assert M.channel_opts["shrink"] == 2 and M.channel_opts["n_per_oct"] == 4 and M.channel_opts["smooth"] == 1
    t = M.classifier[0]
    assert t.feature.dtype == np.uint8 and t.left.tolist() == [1, 2, -1, -1, 5, -1, -1] and t.node_idx.tolist() == [0, 1, 4]
    dst = tmp_path / "copy.pb"
    wb.save(M, str(dst))
    raw = open(dst, "rb").read()
    assert raw[:2] == b"\x78\xda"                                   # zlib level 9 header, like the reference's files
    a, b = model_pb2.Model(), model_pb2.Model()
    a.ParseFromString(zlib.decompress(open(src, "rb").read()))
    b.ParseFromString(zlib.decompress(raw))
    assert a == b and b.channel_opts.func == "waldboost.channels.grad_hist"
    M2 = wb.Model.load(str(dst))
    assert all(np.array_equal(x.threshold, y.threshold) for x, y in zip(M.classifier, M2.classifier)) and M.theta == M2.theta
    bad = tmp_path / "bad.pb"
    bad.write_bytes(b"not a model")
    with pytest.raises(ValueError):
        wb.load(str(bad))


def test_model_container_protocol_and_boxes():
    M = wb.Model((12, 12, 4), OPTS)
    assert not M and len(M) == 0 and M.eval_cost == 0
    tree = S.random_trees((12, 12, 4), 1, 2, np.zeros(4), np.ones(4))[0]
    M.append(tree, -1.5)
    assert M and len(M) == 1 and M[0] == (tree, -1.5) and list(M) == [(tree, -1.5)]
    b = M.get_boxes(np.array([3, 4]), np.array([5, 6]), 0.5)
    assert np.array_equal(b.get(), np.array([[10, 6, 34, 30], [12, 8, 36, 32]], np.float32))
    assert len(M.get_boxes(np.array([]), np.array([]), 0.5)) == 0
    b.set_field("scores", np.array([1.0, 2.0], np.float32))
    c = wb.concatenate([b, b[1]])
    assert len(c) == 3 and c.get_field("scores").tolist() == [1.0, 2.0, 2.0] and c.has_field("scores")
    with pytest.raises(ValueError):
        b.set_field("x", np.zeros(3))


def test_synthetic_recipes_are_seeded():
    assert np.array_equal(S.synthetic_frame(1000, 48, 64), S.synthetic_frame(1000, 48, 64))
    assert S.synthetic_frame(1000, 48, 64).dtype == np.uint8
    l, r = S.full_tree_layout(2)
    assert l.tolist() == [1, 2, -1, -1, 5, -1, -1] and r.tolist() == [4, 3, -1, -1, 6, -1, -1]
    assert S.full_tree_layout(4)[0].size == 31


def test_level_subset_plan_keeps_layout():
    """wbg_plan_create_levels: geometry and offsets of the full pyramid, windows only for the listed levels."""
    spec = wb.channels.resolve_channels(wb.channels.grad_hist)
    full = plan_geometry(480, 640, OPTS, spec, 12, 12)
    ids = [0, 3, 7, 20]
    sub = plan_geometry(480, 640, OPTS, spec, 12, 12, level_ids=ids)
    assert sub.n_levels == full.n_levels and sub.chn_floats == full.chn_floats
    assert [k for k, lv in enumerate(sub.levels) if not lv.skipped] == ids
    assert all(a.chn_off == b.chn_off and a.win_off == b.win_off and a.scale == b.scale for a, b in zip(full.levels, sub.levels))
    assert sub.info.n_loc == sum(full.levels[k].win_rows * full.levels[k].win_cols for k in ids)
    none = plan_geometry(480, 640, OPTS, spec, 12, 12, level_ids=[])          # a rank that got no level
    assert none.info.n_loc == 0 and all(lv.skipped for lv in none.levels)


def test_plan_geometry_property_random_sizes():
    """level sizes must equal the reference's Python-double arithmetic (channels.py:124-131) for arbitrary frames:
    a one-pixel difference in any level would shift every window of that level."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=150, deadline=None)
    @given(st.integers(8, 4500), st.integers(8, 4500), st.integers(1, 12), st.sampled_from([1, 2]))
    def check(Hh, Ww, npo, shrink):
        opts = dict(shrink=shrink, n_per_oct=npo, smooth=1, channels=wb.channels.grad_hist)
        plan = plan_geometry(Hh, Ww, opts, wb.channels.resolve_channels(wb.channels.grad_hist), 12, 12)
        k = 0
        h, w = Hh, Ww
        while not (w < 8 or h < 8):
            for i in range(npo):
                nh, nw = O.level_size(h, w, i, npo, shrink)
                lv = plan.levels[k]
                assert (lv.nh, lv.nw) == (nh, nw) and lv.scale == nw / Ww / shrink
                k += 1
            h, w = h // 2, w // 2
        assert k == plan.n_levels
    check()
