"""Seeded synthetic inputs for tests and benchmarks: frames after the reference's
`fake_data_generator` recipe (reference waldboost/utils.py:81-97) scaled to the frame size, and random-init
cascades of full depth-d trees in sklearn pre-order layout.  NumPy only; no GPU, no oracle."""
import numpy as np

from .training import DTree


def synthetic_frame(seed, H, W):
    """uint8 frame: bright axis-aligned squares on a dark canvas plus uniform noise."""
    rng = np.random.default_rng(seed)
    canvas = np.zeros((H, W), np.float64)
    k = int(rng.integers(0, 1 + int(H * W / 65536 * 2)))
    hi = max(60 * H / 256, 31)
    for _ in range(k):
        side = int(rng.uniform(30, hi))
        y = int(rng.integers(0, max(H - side, 1)))
        x = int(rng.integers(0, max(W - side, 1)))
        canvas[y:y + side, x:x + side] += rng.uniform(0.2, 1.0)
    canvas += rng.random((H, W)) * 0.3 * rng.random()
    return (np.clip(canvas, 0, 1) * 255).astype(np.uint8)


def synthetic_frames(n, H, W, first_seed=1000):
    """frame i uses seed first_seed + i."""
    return np.stack([synthetic_frame(first_seed + i, H, W) for i in range(n)])


def noise_frame(seed, H, W):
    """pure-noise stress frame: maximum gradient energy, fewest rejections."""
    return np.random.default_rng(seed).integers(0, 256, (H, W)).astype(np.uint8)


def full_tree_layout(depth):
    """(left, right) of a full binary tree of the given depth in sklearn pre-order numbering;
    depth 2 -> left=[1,2,-1,-1,5,-1,-1], right=[4,3,-1,-1,6,-1,-1]."""
    left, right = [], []

    def build(d):
        i = len(left)
        left.append(-1)
        right.append(-1)
        if d > 0:
            left[i] = build(d - 1)
            right[i] = build(d - 1)
        return i
    build(depth)
    return np.array(left, np.int8), np.array(right, np.int8)


def random_trees(shape, n_stages, depth, thr_lo, thr_hi, seed=7):
    """n_stages random full trees: internal nodes get a uniform feature (r,c,ch) inside the window and a float32
    threshold uniform in [thr_lo[ch], thr_hi[ch]] (per-channel quantiles of real channel values); leaves get
    prediction float32(N(0, 0.5)), threshold -2 (what sklearn stores), feature (0,0,0)."""
    rng = np.random.default_rng(seed)
    m, n, C = shape
    left, right = full_tree_layout(depth)
    internal = left >= 0
    N = left.size
    trees = []
    for _ in range(n_stages):
        feature = np.zeros((N, 3), np.uint8)
        threshold = np.full(N, -2.0, np.float32)
        prediction = np.zeros(N, np.float32)
        for k in range(N):
            if internal[k]:
                ch = int(rng.integers(0, C))
                feature[k] = (int(rng.integers(0, m)), int(rng.integers(0, n)), ch)
                threshold[k] = np.float32(rng.uniform(thr_lo[ch], thr_hi[ch]))
            else:
                prediction[k] = np.float32(rng.normal(0, 0.5))
        trees.append(DTree(feature, threshold, left, right, prediction))
    return trees


def channel_quantiles(chns, lo=0.10, hi=0.90):
    """per-channel (q_lo, q_hi) of one (u,v,C) channel map."""
    flat = chns.reshape(-1, chns.shape[-1])
    return np.quantile(flat, lo, axis=0), np.quantile(flat, hi, axis=0)


def calibrate_thetas(stage_scores_fn, n_stages, keep_total=1e-4):
    """"wald" rejection profile: theta_t is chosen so that the windows still alive after stage t are (at most) the
    fraction keep_total**((t+1)/T) of the windows that entered stage 0 -- geometric decay to keep_total, i.e. about
    1/(1-keep_total**(1/T)) stages evaluated per window.  Ties (flat image regions give many windows the same score)
    are resolved by rejecting the tied class, never by keeping more than the target; at least one score class is
    kept.  `stage_scores_fn(t, alive_idx)` returns the float32 prediction of stage t for the given window indices
    (any implementation).  Returns float32 thetas."""
    thetas = np.empty(n_stages, np.float32)
    alive, hs, n0 = None, None, 0
    for t in range(n_stages):
        pred = stage_scores_fn(t, alive)
        if hs is None:
            hs = np.zeros(pred.shape, np.float32)
            alive = np.arange(pred.size)
            n0 = pred.size
        hs = hs + pred.astype(np.float32)
        if hs.size == 0:
            thetas[t] = -np.inf
            continue
        target = int(np.floor(n0 * keep_total ** ((t + 1.0) / n_stages)))
        if hs.size <= target:
            thetas[t] = -np.inf
            continue
        srt = np.sort(hs)[::-1]
        th = srt[max(target, 1) - 1]                       # keeps >= target windows if there are ties at th
        if np.count_nonzero(hs >= th) > max(target, 1) and th < srt[0]:
            th = srt[srt > th].min()                       # reject the tied class instead
        thetas[t] = th
        m = hs >= th
        alive, hs = alive[m], hs[m]
    return thetas


def calibrate_thetas_v0(stage_scores_fn, n_stages, keep_total=1e-4):
    """The first calibration recipe: theta_t = the lower quantile of the running score that keeps the fraction
    keep_total**(1/T) of the windows ENTERING stage t.  Kept verbatim because tests/golden/small_model.pb,
    generic_model.pb and configA_model.pb (and the reference outputs recorded with them) were generated with it; ties
    at the quantile keep more windows than the target, which calibrate_thetas fixes for the larger models."""
    keep = keep_total ** (1.0 / n_stages)
    thetas = np.empty(n_stages, np.float32)
    alive, hs = None, None
    for t in range(n_stages):
        pred = stage_scores_fn(t, alive)
        if hs is None:
            hs = np.zeros(pred.shape, np.float32)
            alive = np.arange(pred.size)
        hs = hs + pred.astype(np.float32)
        if hs.size == 0:
            thetas[t] = -np.inf
            continue
        th = np.float32(np.quantile(hs, 1.0 - keep, method="lower"))
        thetas[t] = th
        m = hs >= th
        alive, hs = alive[m], hs[m]
    return thetas
