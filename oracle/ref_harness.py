"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference package from /root/reference.

The reference (`waldboost` 0.2.0) cannot be imported as-is in this image: `skimage`, `bbx` and the
generated `waldboost/model_pb2.py` are absent.  This module pre-seeds `sys.modules`
with three shims so the reference's own source runs unchanged; all arithmetic that matters
(scipy `convolve1d` / `zoom`, the Numba kernels, NumPy gathers) is the real dependency code.

It is used ONLY by `tests/golden/make_golden.py` (fixture generation, in the build container) and by the
optional `tests/test_oracle_vs_reference.py` (skipped when /root/reference is absent, e.g. on the GPU box).
Nothing in the product package, `bench.py` or `smoke()` imports it.

Shims:
  1. `skimage.transform.resize`  -- scikit-image >= 0.19 semantics for the single call at
     waldboost/channels.py:132: scipy.ndimage.zoom(order=1, mode="mirror", grid_mode=True) on a float
     copy of the image (uint8 -> float64, float32 stays float32), then clip to the input's [min, max].
  2. `bbx` / `bbx.boxes`         -- minimal Boxes container + concatenate (model.py:139,147,177,179).
  3. `waldboost.model_pb2`       -- built at runtime from the three messages of waldboost/model.proto:1-23.
"""
import importlib
import os
import sys
import types
import warnings

import numpy as np

REFERENCE_ROOT = os.environ.get("WALDBOOST_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "waldboost"))


# ----------------------------------------------------------------------------------------------- shim 1
def _resize(image, output_shape, order=None, mode="reflect", cval=0, clip=True, preserve_range=False,
            anti_aliasing=None, anti_aliasing_sigma=None):
    import scipy.ndimage as ndi
    assert preserve_range and order == 1 and not anti_aliasing and mode == "reflect"
    img = image if image.dtype.char in "df" else image.astype(np.float64)
    factors = np.divide(img.shape, output_shape)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)
        out = ndi.zoom(img, [1 / f for f in factors], order=1, mode="mirror", cval=cval, grid_mode=True)
    if clip:
        np.clip(out, img.min(), img.max(), out=out)
    return out


# ----------------------------------------------------------------------------------------------- shim 2
class _Boxes:
    def __init__(self, C, **fields):
        self.C = np.asarray(C, np.float32).reshape(-1, 4)
        self.fields = dict()
        for k, v in fields.items():
            self.set_field(k, v)

    def normalized(self, shift=(0, 0), scale=1):
        B = _Boxes((self.C - np.tile(np.asarray(shift, np.float32), 2)) * np.float32(scale))
        B.fields = dict(self.fields)
        return B

    def set_field(self, name, value):
        self.fields[name] = np.asarray(value)

    def get_field(self, name):
        return self.fields[name]

    def has_field(self, name):
        return name in self.fields

    def get(self):
        return self.C

    def __len__(self):
        return self.C.shape[0]

    def __getitem__(self, i):
        B = _Boxes(self.C[i].reshape(-1, 4))
        B.fields = {k: v[i] for k, v in self.fields.items()}
        return B


def _concatenate(boxes, fields=None):
    boxes = list(boxes)
    if not boxes:
        return _Boxes(np.empty((0, 4), "f"))
    if fields is None:
        fields = list(boxes[0].fields.keys())
    B = _Boxes(np.concatenate([b.C for b in boxes], axis=0))
    for f in fields:
        B.set_field(f, np.concatenate([np.atleast_1d(b.get_field(f)) for b in boxes]))
    return B


# ----------------------------------------------------------------------------------------------- shim 3
def _build_model_pb2():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name="waldboost_ref_model.proto", syntax="proto3")

    def msg(name, fields):
        m = fd.message_type.add(name=name)
        for fname, num, ftype, label, tname in fields:
            f = m.field.add(name=fname, number=num, type=ftype, label=label)
            if tname:
                f.type_name = tname
    R, O = F.LABEL_REPEATED, F.LABEL_OPTIONAL
    msg("Model", [("shape", 1, F.TYPE_INT32, R, None), ("channel_opts", 2, F.TYPE_MESSAGE, O, ".ChannelOpts"),
                  ("classifier", 3, F.TYPE_MESSAGE, R, ".DTree"), ("theta", 4, F.TYPE_FLOAT, R, None)])
    msg("ChannelOpts", [("shrink", 1, F.TYPE_INT32, O, None), ("n_per_oct", 2, F.TYPE_INT32, O, None),
                        ("smooth", 3, F.TYPE_INT32, O, None), ("func", 5, F.TYPE_STRING, O, None)])
    msg("DTree", [("feature", 1, F.TYPE_INT32, R, None), ("threshold", 2, F.TYPE_FLOAT, R, None),
                  ("left", 3, F.TYPE_INT32, R, None), ("right", 4, F.TYPE_INT32, R, None),
                  ("prediction", 5, F.TYPE_FLOAT, R, None)])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    mod = types.ModuleType("waldboost.model_pb2")
    for n in ("Model", "ChannelOpts", "DTree"):
        setattr(mod, n, message_factory.GetMessageClass(pool.FindMessageTypeByName(n)))
    return mod


_ref = None


def import_reference():
    """Return the reference `waldboost` package (imported once, unmodified, under the shims)."""
    global _ref
    if _ref is not None:
        return _ref
    if not reference_available():
        raise ImportError(f"reference not found under {REFERENCE_ROOT}")
    if "waldboost" in sys.modules:
        raise RuntimeError("a module named 'waldboost' is already imported")

    sk = types.ModuleType("skimage")
    skt = types.ModuleType("skimage.transform")
    skt.resize = _resize
    sk.transform = skt
    sys.modules.setdefault("skimage", sk)
    sys.modules.setdefault("skimage.transform", skt)

    bbx = types.ModuleType("bbx")
    bbxb = types.ModuleType("bbx.boxes")
    bbx.Boxes = bbxb.Boxes = _Boxes
    bbx.concatenate = _concatenate
    bbx.boxes = bbxb
    sys.modules.setdefault("bbx", bbx)
    sys.modules.setdefault("bbx.boxes", bbxb)

    # numpy >= 1.24 removed np.bool / np.int which the reference still names (model.py:207, __init__.py:128)
    if not hasattr(np, "bool"):
        np.bool = bool
    if not hasattr(np, "int"):
        np.int = int

    sys.modules["waldboost.model_pb2"] = _build_model_pb2()
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            _ref = importlib.import_module("waldboost")
    finally:
        sys.path.remove(REFERENCE_ROOT)
    return _ref
