"""TEST INFRASTRUCTURE ONLY -- CPU restatement (NumPy) of the reference's detect() hot path.

This file is the *oracle* the CUDA path is checked against.  It is never imported by the product package
`waldboost_b200`; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may use it.

Every function cites the reference file:line it restates (paths relative to the reference repository root,
RomanJuranek/waldboost @ 0.2.0).  Arithmetic that the reference delegates to third-party code that is not
in the reference tree is restated from the dependency's documented algorithm and *pinned* against the
dependency itself in `tests/test_oracle_pins.py` (scipy 1.18.1 `ndimage.correlate1d` / `ndimage.zoom`,
numba 0.65 stencil / njit typing, numpy 2.3 NEP-50 promotion) and against the unmodified reference run under
`oracle/ref_harness.py` (`tests/test_oracle_vs_reference.py`, golden fixtures in `tests/golden/`).

Parity status: the reference ships no tests or golden vectors of its own, so parity is
pinned on outputs of the reference itself, generated in the build container by `tests/golden/make_golden.py`.
"""
import math

import numpy as np

F32 = np.float32
F64 = np.float64


# =============================================================================================== channels
def triangle_kernel(n):
    """channels.py:11-13 -- [1..n+1..1]/sum as float32."""
    H = (np.concatenate([np.arange(n + 1), np.arange(n - 1, -1, -1)]) + 1).astype(F32)
    return H / H.sum()


def _reflect_index(i, n):
    """scipy 'reflect' (half-sample symmetric, d c b a | a b c d | d c b a), valid for any offset."""
    i = np.asarray(i)
    if n == 1:
        return np.zeros_like(i)
    p = 2 * n
    i = np.mod(i, p)
    return np.where(i >= n, p - 1 - i, i)


def _correlate1d_sym(x, w, axis):
    """scipy.ndimage correlate1d, mode='reflect', symmetric odd-length weights, float32 in/out.

    Restates the symmetric branch of NI_Correlate1D: accumulate in float64 as
    tmp = x[l]*w[c]; for k = -c..-1: tmp += (x[l+k] + x[l-k]) * w[c+k]; result cast to the input dtype.
    Called from channels.py:19-20 (H=[1,2,1]) and channels.py:25-26 (triangle kernel).
    """
    x = np.moveaxis(np.asarray(x, F32), axis, 0)
    n = x.shape[0]
    c = len(w) // 2
    w = np.asarray(w, F64)
    idx = np.arange(n)
    xd = x.astype(F64)
    acc = xd * w[c]
    for k in range(-c, 0):
        acc = acc + (xd[_reflect_index(idx + k, n)] + xd[_reflect_index(idx - k, n)]) * w[c + k]
    return np.moveaxis(acc.astype(F32), 0, axis)


def _convolve1d_diff(x, axis):
    """convolve1d(x, [-1,0,1], axis) (channels.py:18-20): convolution flips the kernel, so the result is
    x[l-1] - x[l+1] with 'reflect' borders (anti-symmetric branch of NI_Correlate1D, float64 accumulate)."""
    x = np.moveaxis(np.asarray(x, F32), axis, 0)
    n = x.shape[0]
    idx = np.arange(n)
    xd = x.astype(F64)
    acc = xd * 0.0 + (xd[_reflect_index(idx - 1, n)] - xd[_reflect_index(idx + 1, n)]) * 1.0
    return np.moveaxis(acc.astype(F32), 0, axis)


def gradients(image):
    """channels.py:16-21 -- gy = D_rows(H_cols(I)), gx = D_cols(H_rows(I)); negated Sobel."""
    image = np.asarray(image, F32)
    H = np.array([1, 2, 1], F32)
    gy = _convolve1d_diff(_correlate1d_sym(image, H, 1), 0)
    gx = _convolve1d_diff(_correlate1d_sym(image, H, 0), 1)
    return gx, gy


def separable_convolve(image, k0):
    """channels.py:24-27 -- symmetric kernel along axis 0 then axis 1."""
    return _correlate1d_sym(_correlate1d_sym(image, k0, 0), k0, 1)


def grad_mag(image, norm=5, eps=1e-3):
    """channels.py:30-37."""
    gx, gy = gradients(np.asarray(image).astype(F32))
    mag = np.sqrt(gx * gx + gy * gy)              # float32 throughout
    if norm is not None and norm > 1:
        nrm = separable_convolve(mag, triangle_kernel(norm))
        mag = mag / (nrm + F32(eps))              # python-float eps is weak under NEP 50 -> float32
    return mag[..., None]


def grad_hist(image, n_bins=4, full=False, bias=0):
    """channels.py:40-52.  cos/sin are float64 scalars, so under NumPy 2 (NEP 50) gx*c - gy*s is evaluated
    in float64 (two rounded products, one rounded difference) and rounded once to float32 on store."""
    gx, gy = gradients(np.asarray(image).astype(F32))
    max_theta = 2 * np.pi if full else np.pi
    theta = np.linspace(0, max_theta, n_bins + 1)
    cs = np.cos(theta[:-1])
    sn = np.sin(theta[:-1])
    chns = np.empty(gx.shape + (n_bins,), F32)
    gxd, gyd = gx.astype(F64), gy.astype(F64)
    for i in range(n_bins):
        chns[..., i] = gxd * cs[i] - gyd * sn[i]
    val = np.fmax(np.abs(chns) - F32(bias), F32(0))
    return (np.sign(chns) * val) if full else val


def _fpga_grad(arr):
    """fpga/channels.py:5-27 -- integer Sobel stencils; Numba's stencil leaves the 1-pixel border at cval = 0 and
    widens uint8 operands to int64; the results are stored as int32 (fpga/channels.py:40-43)."""
    a = np.asarray(arr).astype(np.int64)
    dx = np.zeros(a.shape, np.int64)
    dy = np.zeros(a.shape, np.int64)
    if a.shape[0] >= 3 and a.shape[1] >= 3:
        c = lambda dr, dc: a[1 + dr:a.shape[0] - 1 + dr, 1 + dc:a.shape[1] - 1 + dc]
        dx[1:-1, 1:-1] = -(c(-1, -1) + 2 * c(0, -1) + c(1, -1)) + c(-1, 1) + 2 * c(0, 1) + c(1, 1)
        dy[1:-1, 1:-1] = -(c(-1, -1) + 2 * c(-1, 0) + c(-1, 1)) + c(1, -1) + 2 * c(1, 0) + c(1, 1)
    return dx.astype(np.int32), dy.astype(np.int32)


def grad_hist_4_u1(image):
    """fpga/channels.py:29-52 -- dx, 0.5dx-0.5dy, dy, 0.5dx+0.5dy (float64, truncated toward zero on the int32 store),
    then min(|y| // 4, 255) as uint8."""
    dx, dy = _fpga_grad(image)
    y = np.empty(dx.shape + (4,), np.int32)
    y[..., 0] = dx
    y[..., 1] = np.trunc(0.5 * dx - 0.5 * dy)
    y[..., 2] = dy
    y[..., 3] = np.trunc(0.5 * dx + 0.5 * dy)
    return np.fmin(np.abs(y) // 4, 255).astype(np.uint8)


def grad_mag_u1(image):
    """fpga/channels.py:55-66 -- max(|dx|, |dy|) // 4 clamped to 255, uint8."""
    dx, dy = _fpga_grad(image)
    y = np.maximum(np.abs(dx), np.abs(dy))[..., None]
    return np.fmin(y // 4, 255).astype(np.uint8)


def grad_mag_hist(image, n_bins=9, norm=5, eps=1e-3):
    """Composite 1+n_bins channel function used by BASELINE config C: the reference has no
    such function; it is defined as the concatenation of the two reference functions on the same image."""
    return np.concatenate([grad_mag(image, norm, eps), grad_hist(image, n_bins)], axis=-1)


def avg_pool_2(arr):
    """channels.py:55-64.  Numba typing: uint8 operands are widened to int64 before the adds (no wrap),
    '/4' is a float64 true division and astype(uint8) truncates; float32 stays float32 with the add order
    ((a00 + a10) + a01) + a11."""
    u, v = arr.shape[:2]
    ul, vl = u - u % 2, v - v % 2
    a00, a10 = arr[0:ul:2, 0:vl:2, ...], arr[1:ul:2, 0:vl:2, ...]
    a01, a11 = arr[0:ul:2, 1:vl:2, ...], arr[1:ul:2, 1:vl:2, ...]
    if arr.dtype.kind in "ui":
        s = ((a00.astype(np.int64) + a10) + a01) + a11
        return (s / 4).astype(arr.dtype)
    return ((((a00 + a10) + a01) + a11) / arr.dtype.type(4)).astype(arr.dtype)


def max_pool_2(arr):
    """channels.py:67-75 (defined by the reference, never called by the pyramid)."""
    u, v = arr.shape[:2]
    ul, vl = u - u % 2, v - v % 2
    m0 = np.fmax(arr[0:ul:2, 0:vl:2, ...], arr[1:ul:2, 0:vl:2, ...])
    m1 = np.fmax(arr[0:ul:2, 1:vl:2, ...], arr[1:ul:2, 1:vl:2, ...])
    return np.fmax(m0, m1)


def smooth_image_3d(arr):
    """channels.py:78-90.  3x3 [1 2 1]x[1 2 1]: Numba types int*float32 as float64, so the nine terms are
    accumulated in float64 in source order, divided by 16 and rounded to float32; the stencil leaves the
    1-pixel border at cval=0."""
    if np.asarray(arr).dtype.kind in "ui":
        # integer channel maps (fpga channels): int64 sums, true division, truncating store into the input dtype
        a = np.asarray(arr).astype(np.int64)
        out = np.zeros(a.shape, np.int64)
        if a.shape[0] >= 3 and a.shape[1] >= 3:
            c = lambda dr, dc: a[1 + dr:a.shape[0] - 1 + dr, 1 + dc:a.shape[1] - 1 + dc]
            out[1:-1, 1:-1] = (c(-1, -1) + 2 * c(-1, 0) + c(-1, 1) + 2 * c(0, -1) + 4 * c(0, 0) + 2 * c(0, 1)
                               + c(1, -1) + 2 * c(1, 0) + c(1, 1))
        return (out / 16).astype(np.asarray(arr).dtype)
    a = np.asarray(arr, F32).astype(F64)
    out = np.zeros(a.shape, F64)
    if a.shape[0] >= 3 and a.shape[1] >= 3:
        c = lambda dr, dc: a[1 + dr:a.shape[0] - 1 + dr, 1 + dc:a.shape[1] - 1 + dc]
        v = c(-1, -1) + 2 * c(-1, 0)
        v = v + c(-1, 1)
        v = v + 2 * c(0, -1)
        v = v + 4 * c(0, 0)
        v = v + 2 * c(0, 1)
        v = v + c(1, -1)
        v = v + 2 * c(1, 0)
        v = v + c(1, 1)
        out[1:-1, 1:-1] = v
    return (out / 16).astype(F32)


# =============================================================================================== pyramid
def image_octaves(image):
    """channels.py:93-101 -- the size test happens before the yield."""
    base = image.copy()
    while True:
        h, w = base.shape[:2]
        if w < 8 or h < 8:
            break
        yield base
        base = avg_pool_2(base)


def level_size(h, w, i, n_per_oct, shrink):
    """channels.py:124-130 -- Python double arithmetic, must not be re-associated."""
    factor = 2 ** (-1 / n_per_oct)
    s = factor ** i
    return int((h * s) / shrink) * shrink, int((w * s) / shrink) * shrink   # (nh, nw)


def _axis_taps(n_in, n_out):
    """scipy NI_ZoomShift, order=1, grid_mode=True: cc = (j + 0.5) * (n_in / n_out) - 0.5 evaluated in three
    float64 steps; start = floor(cc); weights (1 - t, t) with t = cc - floor(cc)."""
    zoom = np.divide(np.int64(n_in), np.int64(n_out))           # float64
    cc = np.arange(n_out, dtype=F64)
    cc = cc + 0.5
    cc = cc * zoom
    cc = cc - 0.5
    fl = np.floor(cc)
    t = cc - fl
    i0 = fl.astype(np.int64)
    i1 = i0 + 1
    # 'mirror' extension (d c b | a b c d | c b a); for down-scaling only i1 == n_in with weight 0 occurs
    i0 = np.clip(i0, 0, n_in - 1)
    i1 = np.where(i1 > n_in - 1, np.maximum(2 * (n_in - 1) - i1, 0), i1)
    return i0, i1, 1.0 - t, t


def resize_bilinear(base, nh, nw):
    """channels.py:132 -- skimage.transform.resize(base, (nh, nw), preserve_range=True, order=1,
    anti_aliasing=False).astype(base.dtype).

    skimage >= 0.19 forwards to scipy.ndimage.zoom(order=1, mode='mirror', grid_mode=True) on a float copy
    (uint8 -> float64, float32 stays float32) and clips to the input's [min, max]; scipy returns the input
    unchanged when both zoom factors are exactly 1.  The interpolation sum runs over the 2x2 footprint in
    row-major order, each term being (value * w_row) * w_col, accumulated in float64 starting from 0.
    """
    h, w = base.shape
    dtype = base.dtype
    img = base if dtype.char in "df" else base.astype(F64)
    if nh == h and nw == w:
        out = img.copy()
    else:
        r0, r1, wr0, wr1 = _axis_taps(h, nh)
        c0, c1, wc0, wc1 = _axis_taps(w, nw)
        v = img.astype(F64)
        wr0, wr1 = wr0[:, None], wr1[:, None]
        wc0, wc1 = wc0[None, :], wc1[None, :]
        t = (v[r0][:, c0] * wr0) * wc0
        t = t + (v[r0][:, c1] * wr0) * wc1
        t = t + (v[r1][:, c0] * wr1) * wc0
        t = t + (v[r1][:, c1] * wr1) * wc1
        out = t.astype(img.dtype)
    out = np.clip(out, img.min(), img.max())
    return out.astype(dtype)        # uint8: truncation toward zero


def channel_pyramid(image, channel_opts, levels=None):
    """channels.py:111-146.  `channel_opts["channels"]` is a callable im -> (h, w, C) float32.
    `levels` (not in the reference): if given, only the pyramid levels with these indices are computed and yielded
    (every level depends only on the original image) -- used to spread one frame over several CPU workers."""
    if not isinstance(image, np.ndarray):
        raise TypeError("Image must be numpy array")
    if image.ndim != 2:
        raise ValueError("Image must have 2 dimensions")
    shrink = channel_opts["shrink"]
    n_per_oct = channel_opts["n_per_oct"]
    smooth = channel_opts["smooth"]
    channels = channel_opts["channels"]
    assert shrink in [1, 2], "Shrink factor must be integer 1 <= shrink <= 2"
    index = -1
    for base in image_octaves(image):
        h, w = base.shape[:2]
        for i in range(n_per_oct):
            index += 1
            if levels is not None and index not in levels:
                continue
            nh, nw = level_size(h, w, i, n_per_oct, shrink)
            real_scale = nw / image.shape[1]
            im = resize_bilinear(base, nh, nw)
            chns = channels(im)
            if shrink == 2:
                chns = avg_pool_2(chns)
            if smooth == 1:
                chns = smooth_image_3d(chns)
            yield np.atleast_3d(chns), real_scale / shrink


# =============================================================================================== cascade
class DTree:
    """training.py:23-31 -- array form of one decision-tree stage (inference half only)."""

    def __init__(self, feature, threshold, left, right, prediction):
        self.feature = np.array([f if f is not None else [0, 0, 0] for f in feature], np.uint8).reshape(-1, 3)
        self.threshold = np.array(threshold, F32)
        self.left = np.array(left, np.int8)
        self.right = np.array(right, np.int8)
        self.prediction = np.array(prediction, F32)
        self.node = self.left >= 0
        self.node_idx = np.flatnonzero(self.node)

    def leaf_on_image(self, X, rs, cs):
        """training.py:84-95 -- breadth-wise traversal; returns the final node (leaf) index per window."""
        node = np.zeros(rs.size, np.int32)
        idx_in_node = {0: np.arange(rs.size)}
        for n in self.node_idx:
            r, c, ch = (int(t) for t in self.feature[n])
            lnode, rnode = int(self.left[n]), int(self.right[n])
            idx = idx_in_node[n]
            go_left = X[rs[idx] + r, cs[idx] + c, ch] <= self.threshold[n]
            node[idx] = np.where(go_left, lnode, rnode)
            idx_in_node[lnode] = idx[go_left]
            idx_in_node[rnode] = idx[~go_left]
        return node

    def predict_on_image(self, X, rs, cs):
        """training.py:84-96."""
        return self.prediction[self.leaf_on_image(X, rs, cs)]

    def apply(self, X):
        """training.py:73-81 -- sample mode, X is (N, m, n, C)."""
        node = np.zeros(X.shape[0], np.int32)
        for n in self.node_idx:
            r, c, ch = (int(t) for t in self.feature[n])
            idx = np.flatnonzero(node == n)
            go_left = X[idx, r, c, ch] <= self.threshold[n]
            node[idx] = np.where(go_left, self.left[n], self.right[n])
        return node

    def predict(self, X):
        """training.py:82-83."""
        return self.prediction[self.apply(X)]


class Cascade:
    """model.py:32-283 restricted to the inference path: shape, channel_opts, stages, thetas, stats."""

    def __init__(self, shape, channel_opts):
        self.shape = tuple(shape)
        self.channel_opts = channel_opts
        self.classifier = []
        self.theta = []
        self.n_loc = 0
        self.n_weak = 0

    def append(self, weak, theta):
        self.classifier.append(weak)
        self.theta.append(theta)

    def __len__(self):
        return len(self.classifier)

    @property
    def eval_cost(self):
        """model.py:69-84."""
        return self.n_weak / self.n_loc if self.n_loc > 0 else 0

    def reset(self):
        self.n_loc = 0
        self.n_weak = 0

    def channels(self, image, levels=None):
        """model.py:95-103."""
        yield from channel_pyramid(image, self.channel_opts, levels)

    def predict_on_image(self, X, trace=None):
        """model.py:216-259.  Window grid is (u-m) x (v-n) (model.py:243); scores accumulate in float32 in stage
        order; a stage with theta == -inf does not filter.  `trace`, if a list, receives the number of windows
        entering each evaluated stage."""
        u, v, ch_image = X.shape
        m, n, ch_cls = self.shape
        assert ch_image == ch_cls, f"Invalid number of channels. Expected {ch_cls} given {ch_image}."
        rs, cs = np.indices((max(u - m, 0), max(v - n, 0)))
        rs = rs.flatten()
        cs = cs.flatten()
        hs = np.zeros(rs.shape, F32)
        self.n_loc += hs.size
        for weak, theta in zip(self.classifier, self.theta):
            if not rs.size:
                break
            hs += weak.predict_on_image(X, rs, cs)
            self.n_weak += hs.size
            if trace is not None:
                trace.append(hs.size)
            if theta == -np.inf:
                continue
            mask = hs >= theta
            rs, cs, hs = rs[mask], cs[mask], hs[mask]
        return rs, cs, hs

    def get_boxes(self, r, c, scale):
        """model.py:136-147 -- [c, r, c+n, r+m] as float32, multiplied by float32(1/scale)
        (bbx.Boxes.normalized(scale=k) is taken to be coords * k)."""
        if r.size == 0:
            return np.empty((0, 4), F32)
        m, n = self.shape[:2]
        x1 = c.reshape(-1, 1)
        y1 = r.reshape(-1, 1)
        rects = np.concatenate([x1, y1, x1 + n, y1 + m], axis=1).astype(F32)
        return rects * F32(1.0 / scale)

    def scan_channels(self, image):
        """model.py:105-134."""
        for chns, scale in self.channels(image):
            yield chns, scale, self.predict_on_image(chns)

    def detect(self, image, levels=None):
        """model.py:149-179 -> (boxes [K,4] f32, scores [K] f32, level [K] i32) in (level, r, c) order.
        `levels`: optional sorted list of pyramid level indices to restrict the scan to (see channel_pyramid)."""
        B, S, L = [np.empty((0, 4), F32)], [np.empty(0, F32)], [np.empty(0, np.int32)]
        ids = range(10 ** 9) if levels is None else sorted(levels)
        for lvl, (chns, scale) in zip(ids, self.channels(image, levels)):
            r, c, h = self.predict_on_image(chns)
            B.append(self.get_boxes(r, c, scale))
            S.append(h)
            L.append(np.full(r.size, lvl, np.int32))
        return np.concatenate(B, axis=0), np.concatenate(S), np.concatenate(L)

    def predict(self, X):
        """model.py:181-214 -- sample mode."""
        n = X.shape[0]
        assert tuple(X.shape[1:]) == tuple(self.shape)
        H = np.zeros(n, F32)
        mask = np.ones(n, bool)
        for weak, theta in zip(self.classifier, self.theta):
            H[mask] += weak.predict(X[mask, ...])
            if theta == -np.inf:
                continue
            mask = np.logical_and(mask, H >= theta)
        H[~mask] = -np.inf
        return H, mask


def detect_multi(image, cascades, channel_opts=None, response_scale=None):
    """waldboost.detect(image, *models, channel_opts=, response_scale=) -- __init__.py:75-130: one pyramid shared by all
    models; per level every model's predict_on_image; scores times response_scale[k]; label = model index.
    -> (boxes [K,4] f32, scores [K] f32, label [K] i64) in the order of the reference's loops (level, model, r, c)."""
    channel_opts = channel_opts or cascades[0].channel_opts
    if response_scale is None:
        response_scale = [1] * len(cascades)
    response_scale = np.array(response_scale, "f")
    if response_scale.size != len(cascades):
        raise ValueError("Wrong response_scale parameter")
    B, S, L = [np.empty((0, 4), F32)], [np.empty(0, F32)], [np.empty(0, np.int64)]
    for chns, scale in channel_pyramid(image, channel_opts):
        for k, Cs in enumerate(cascades):
            r, c, h = Cs.predict_on_image(chns)
            if r.size > 0:                                   # __init__.py:125
                B.append(Cs.get_boxes(r, c, scale))
                S.append(h * response_scale[k])
                L.append(np.full(r.size, k, np.int64))
    return np.concatenate(B, axis=0), np.concatenate(S), np.concatenate(L)


def gather_samples(chns, rs, cs, shape):
    """samples.py:14-43."""
    if rs.size != cs.size:
        raise ValueError("Sizes of 'rs' and 'cs' must match")
    m, n, _ = shape
    if rs.size == 0:
        return np.empty((0,) + tuple(shape), dtype=chns.dtype)
    return np.array([chns[r:r + m, c:c + n, ...] for r, c in zip(rs, cs)])


# =============================================================================================== test glue
def oracle_channel_fn(fn):
    """Map a product channel function (waldboost_b200.channels.* / waldboost_b200.fpga.*, possibly wrapped in
    functools.partial with keyword arguments) to the oracle function of the same name.  Name based, so that this
    module never imports the product package."""
    import functools
    kw = {}
    base = fn
    while isinstance(base, functools.partial):
        kw = {**base.keywords, **kw}
        base = base.func
    table = {"grad_hist": grad_hist, "grad_mag": grad_mag, "grad_mag_hist": grad_mag_hist,
             "grad_hist_4_u1": grad_hist_4_u1, "grad_mag_u1": grad_mag_u1}
    target = table[base.__name__]
    return functools.partial(target, **kw) if kw else target


def cascade_from_model(model):
    """product Model (duck typed: shape, channel_opts, classifier with DTree arrays, theta) -> oracle Cascade."""
    opts = dict(model.channel_opts, channels=oracle_channel_fn(model.channel_opts["channels"])) if model.channel_opts else None
    Cs = Cascade(model.shape, opts)
    for w, th in zip(model.classifier, model.theta):
        Cs.append(DTree([tuple(f) for f in w.feature], w.threshold, w.left, w.right, w.prediction), th)
    return Cs
