"""Channel features on the GPU -- same names as the reference module waldboost/channels.py so that model files
whose `channel_opts.func` is "waldboost.channels.grad_hist" resolve (reference waldboost/model.py:302,316).

The functions below are *markers with a GPU implementation*: used as `channel_opts["channels"]` they select the
fused CUDA channel kernel; called directly on an image they run that kernel on the single image.  Arbitrary Python
callables are rejected -- there is no CPU path to run them on.
"""
import functools

import numpy as np

from . import _native as N

__all__ = ["grad_hist", "grad_mag", "grad_mag_hist", "channel_pyramid", "avg_pool_2", "max_pool_2",
           "smooth_image_3d", "triangle_kernel", "gradients", "separable_convolve", "resolve_channels"]


def triangle_kernel(n):
    """reference channels.py:11-13."""
    H = (np.r_[:n + 1, n - 1:-1:-1] + 1).astype("f")
    return H / H.sum()


def _image_f32(image, what):
    a = np.asarray(image)
    if a.ndim != 2 or a.dtype != np.float32:
        raise TypeError(f"{what} on the GPU takes a 2-D float32 image")
    return a


def gradients(image):
    """(gx, gy) = negated Sobel gradients: H = [1,2,1] along one axis, D = [-1,0,1] convolved along the other, 'reflect'
    borders (reference channels.py:16-21); 2-D float32 image -> two float32 arrays."""
    from .engine import get_engine
    return get_engine().gradients(_image_f32(image, "gradients"))


def separable_convolve(image, k0, k1=None):
    """convolve1d(image, k0, axis=0) followed by convolve1d(., k1 or k0, axis=1) (reference channels.py:24-27).  The
    kernels must be symmetric with an odd length <= 63 -- what the reference itself passes (triangle_kernel); other
    kernels take a different accumulation order in scipy and are rejected (ValueError)."""
    from .engine import get_engine
    for k in (k0,) if k1 is None else (k0, k1):
        k = np.asarray(k, np.float32)
        if k.ndim != 1 or k.size % 2 == 0 or k.size > 63 or not np.array_equal(k, k[::-1]):
            raise ValueError("separable_convolve on the GPU takes symmetric 1-D kernels of odd length <= 63")
    return get_engine().separable_convolve(_image_f32(image, "separable_convolve"), k0, k1)


# ----------------------------------------------------------------------------------------------- channel functions
def _direct(kind_fn, image, **kw):
    from .engine import get_engine
    if not isinstance(image, np.ndarray) or image.ndim != 2:
        raise ValueError("channel functions take a 2-D numpy image")
    spec = resolve_channels(functools.partial(kind_fn, **kw) if kw else kind_fn)
    if spec.get("integer"):
        img = np.ascontiguousarray(image)                      # integer channels work on the uint8 image itself
    else:
        img = np.ascontiguousarray(image).astype(np.float32)   # reference: image.astype("f") (channels.py:31,41)
    opts = dict(shrink=1, n_per_oct=1, smooth=0, channels=None)
    levels = get_engine().channel_levels(img, opts, spec, max_levels=1)
    return levels[0][0]


def grad_hist(image, n_bins=4, full=False, bias=0):
    """Oriented gradient channels |gx cos(t) - gy sin(t)| (reference channels.py:40-52) -> (h, w, n_bins) float32."""
    return _direct(grad_hist, image, n_bins=n_bins, full=full, bias=bias)


def grad_mag(image, norm=5, eps=1e-3):
    """Triangle-normalised gradient magnitude (reference channels.py:30-37) -> (h, w, 1) float32."""
    return _direct(grad_mag, image, norm=norm, eps=eps)


def grad_mag_hist(image, n_bins=9, norm=5, eps=1e-3):
    """concat(grad_mag(image, norm, eps), grad_hist(image, n_bins)) -> (h, w, 1+n_bins); the 10-channel feature of
    BASELINE config C.  Not in the reference; defined as the concatenation of its two functions."""
    return _direct(grad_mag_hist, image, n_bins=n_bins, norm=norm, eps=eps)


_DEFAULTS = {
    grad_hist: dict(kind=N.WBG_CH_GRAD_HIST, n_bins=4, full=False, bias=0, norm=0, eps=0.0),
    grad_mag: dict(kind=N.WBG_CH_GRAD_MAG, n_bins=0, full=False, bias=0, norm=5, eps=1e-3),
    grad_mag_hist: dict(kind=N.WBG_CH_GRAD_MAG_HIST, n_bins=9, full=False, bias=0, norm=5, eps=1e-3),
}


_EXTRA = {}          # functions registered by sub-packages (waldboost_b200.fpga): fn -> (spec defaults, channel count)


def register_channel_function(fn, defaults, channels, integer=False):
    _DEFAULTS[fn] = dict(defaults)
    _EXTRA[fn] = dict(channels=channels, integer=integer)


def resolve_channels(fn):
    """Map `channel_opts["channels"]` to the kernel variant: one of the functions above, optionally wrapped in
    functools.partial with keyword arguments.  Anything else raises -- no CPU fallback."""
    kw = {}
    base = fn
    while isinstance(base, functools.partial):
        if base.args:
            raise TypeError("channel function partials may only bind keyword arguments")
        kw = {**base.keywords, **kw}
        base = base.func
    if base not in _DEFAULTS:
        raise TypeError(f"channel function {fn!r} has no CUDA implementation; supported: "
                        "waldboost_b200.channels.grad_hist, grad_mag, grad_mag_hist, waldboost_b200.fpga.grad_hist_4_u1, "
                        "grad_mag_u1 (optionally functools.partial with keyword arguments)")
    if base in _EXTRA and kw:
        raise TypeError(f"{base.__name__}() takes no keyword arguments")
    spec = dict(_DEFAULTS[base])
    for k, v in kw.items():
        if k not in spec or (base is grad_hist and k in ("norm", "eps")) or (base is grad_mag and k in ("n_bins", "full", "bias")):
            raise TypeError(f"{base.__name__}() got an unexpected keyword argument {k!r}")
        spec[k] = v
    if spec["norm"] is None:
        spec["norm"] = 0
    spec["name"] = base.__name__
    spec["integer"] = bool(_EXTRA.get(base, {}).get("integer"))
    if base in _EXTRA:
        spec["channels"] = _EXTRA[base]["channels"]
    return spec


def channel_count(spec):
    if "channels" in spec:
        return spec["channels"]
    return {N.WBG_CH_GRAD_HIST: spec["n_bins"], N.WBG_CH_GRAD_MAG: 1, N.WBG_CH_GRAD_MAG_HIST: 1 + spec["n_bins"]}[spec["kind"]]


# ----------------------------------------------------------------------------------------------- pyramid
def _validate_image(image):
    """reference channels.py:104-108."""
    if not isinstance(image, np.ndarray):
        raise TypeError("Image must be numpy array")
    if image.ndim != 2:
        raise ValueError("Image must have 2 dimensions")


def channel_pyramid(image, channel_opts):
    """Generator of (chns (u,v,C) float32, scale) per pyramid level (reference channels.py:111-146), computed by the
    fused CUDA kernels; the arrays are NumPy copies on the host like the reference's."""
    from .engine import get_engine
    _validate_image(image)
    assert channel_opts["shrink"] in [1, 2], "Shrink factor must be integer 1 <= shrink <= 2"
    spec = resolve_channels(channel_opts["channels"])
    yield from get_engine().channel_levels(image, channel_opts, spec)


# ----------------------------------------------------------------------------------------------- primitives
def _single(fn_name, arr, out_shape):
    from .engine import get_engine
    return get_engine().map_primitive(fn_name, arr, out_shape)


def avg_pool_2(arr):
    """2x2 mean, odd trailing row/column dropped (reference channels.py:55-64); float32 channel maps (u,v[,C])."""
    a = np.asarray(arr)
    if a.dtype != np.float32:
        raise TypeError("avg_pool_2 on the GPU takes float32 channel maps")
    return _single("wbg_avg_pool_2", a, (a.shape[0] // 2, a.shape[1] // 2) + a.shape[2:])


def max_pool_2(arr):
    """2x2 max (reference channels.py:67-75)."""
    a = np.asarray(arr)
    if a.dtype != np.float32:
        raise TypeError("max_pool_2 on the GPU takes float32 channel maps")
    return _single("wbg_max_pool_2", a, (a.shape[0] // 2, a.shape[1] // 2) + a.shape[2:])


def smooth_image_3d(arr):
    """3x3 [1 2 1]x[1 2 1]/16 with a zero border ring (reference channels.py:78-90); (u,v,C) float32."""
    a = np.asarray(arr)
    if a.dtype != np.float32 or a.ndim != 3:
        raise TypeError("smooth_image_3d takes a float32 (u,v,C) array")
    return _single("wbg_smooth_image_3d", a, a.shape)
