"""Shared test helpers: the same cascade as a product Model (GPU) and as an oracle Cascade (CPU)."""
import os

import numpy as np

import wb_oracle as O
import waldboost_b200 as wb
from waldboost_b200 import synthetic as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

def oracle_opts(channel_opts):
    """product channel_opts (functions of waldboost_b200.channels / .fpga, maybe partial) -> oracle channel_opts."""
    return dict(channel_opts, channels=O.oracle_channel_fn(channel_opts["channels"]))


def oracle_cascade(model):
    """wb.Model -> oracle Cascade with identical arrays."""
    return O.cascade_from_model(model)


def make_model(shape, channel_opts, n_stages, depth, frame, seed=7, keep_total=None, calib_levels=1):
    """random cascade with thresholds drawn from the channel quantiles of `frame` (oracle pyramid); thetas are
    -inf (dense) or calibrated on the oracle ('wald' profile) when keep_total is given."""
    levels = list(O.channel_pyramid(frame, oracle_opts(channel_opts)))
    lo, hi = S.channel_quantiles(levels[0][0])
    trees = S.random_trees(shape, n_stages, depth, lo, hi, seed=seed)
    M = wb.Model(shape, channel_opts)
    for t in trees:
        M.append(t, -np.inf)
    if keep_total is not None:
        M.theta = [float(x) for x in calibrate_on_oracle(M, [c for c, _ in levels[:calib_levels]], keep_total)]
    return M


def calibrate_on_oracle(M, maps, keep_total):
    Cs = oracle_cascade(M)
    m, n, _ = M.shape
    mp, R, Cc = [], [], []
    for k, X in enumerate(maps):
        u, v, _ = X.shape
        rs, cs = np.indices((max(u - m, 0), max(v - n, 0)))
        mp.append(np.full(rs.size, k)); R.append(rs.ravel()); Cc.append(cs.ravel())
    mp, R, Cc = np.concatenate(mp), np.concatenate(R), np.concatenate(Cc)

    def stage(t, alive):
        idx = np.arange(mp.size) if alive is None else alive
        out = np.empty(idx.size, np.float32)
        for k, X in enumerate(maps):
            sel = mp[idx] == k
            out[sel] = Cs.classifier[t].predict_on_image(X, R[idx][sel], Cc[idx][sel])
        return out
    return S.calibrate_thetas(stage, len(M), keep_total)


def channels_close(got, ref, rtol=1e-5):
    """parity gate of SURVEY.md 8d: |d| <= rtol * max(|ref|, 1); returns (ok, n_bad, max_err_ratio, n_not_bitexact)."""
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert got.dtype == np.float32
    err = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    tol = rtol * np.maximum(np.abs(ref.astype(np.float64)), 1.0)
    bad = int((err > tol).sum())
    return bad == 0, bad, float((err / tol).max()) if err.size else 0.0, int((got != ref).sum())
