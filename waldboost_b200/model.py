"""Detection model with the reference's API (reference waldboost/model.py:32-344), evaluated on the GPU.

`Model.detect`, `scan_channels`, `channels` and `predict_on_image` keep the reference's signatures, return types
(NumPy on the host, `Boxes` with a "scores" field) and error behaviour, but every pixel and every window is
computed by the CUDA library behind include/wbg.h.  `detect_batch` is the batched entry point the reference
lacks (it loops over images in Python).
"""
import zlib
from importlib import import_module

import numpy as np
from google.protobuf.message import DecodeError

from . import channels as _channels
from . import model_pb2
from .boxes import Boxes
from .channels import _validate_image, resolve_channels
from .engine import ModelHandle, get_engine
from .training import DTree

# names under which channel functions are stored in .pb files (model.py:302).  Functions that exist in the
# reference are written with the reference's module path so either implementation can load the file.
_REFERENCE_NAMES = {("channels", "grad_hist"), ("channels", "grad_mag"),
                    ("fpga.channels", "grad_hist_4_u1"), ("fpga.channels", "grad_mag_u1")}


def symbol_name(s):
    """reference model.py:23-24, with reference-compatible names for the functions both packages have."""
    mod = getattr(s, "__module__", "") or ""
    if mod.startswith(__package__ + ".") and (mod[len(__package__) + 1:], s.__name__) in _REFERENCE_NAMES:
        return "waldboost." + mod[len(__package__) + 1:] + "." + s.__name__
    return s.__module__ + "." + s.__qualname__


def symbol_from_name(name: str):
    """reference model.py:27-29; "waldboost.<x>" resolves inside this package (the reference is not importable)."""
    module, _, symbol = name.rpartition(".")
    if module == "waldboost" or module.startswith("waldboost."):
        module = __package__ + module[len("waldboost"):]
    if not module:
        raise ValueError(f"channel function {name!r} must be a dotted module path")
    m = import_module(module)
    obj = m
    for part in symbol.split("."):
        obj = getattr(obj, part)
    return obj


class Model:
    """Detection model: window shape (m, n, C), channel options and the cascade (stages + rejection thresholds).

        model = Model(shape, channel_opts)          # empty
        model = Model.load("model.pb")              # from the reference's protobuf format
        boxes = model.detect(image)                 # Boxes with field "scores"
    """

    def __init__(self, shape, channel_opts):
        self.shape = shape
        self.channel_opts = channel_opts
        self.classifier = []
        self.theta = []
        self.reset()
        self._handle = None
        self._handle_key = None
        self._handle_trees = None

    # ------------------------------------------------------------------------------------------- stats
    @property
    def eval_cost(self):
        """Average number of weak classifiers evaluated per location (reference model.py:69-84)."""
        return self.n_weak / self.n_loc if self.n_loc > 0 else 0

    def reset(self):
        """reference model.py:86-89."""
        self.n_loc = 0
        self.n_weak = 0

    # ------------------------------------------------------------------------------------------- container protocol
    def __getitem__(self, i):
        return self.classifier[i], self.theta[i]

    def __len__(self):
        return len(self.classifier)

    def __bool__(self):
        return bool(self.classifier)

    def __iter__(self):
        yield from zip(self.classifier, self.theta)

    def append(self, weak, theta):
        """reference model.py:272-283."""
        self.classifier.append(weak)
        self.theta.append(theta)

    # ------------------------------------------------------------------------------------------- device state
    def invalidate(self):
        """Force a re-upload of the cascade (needed only after mutating a DTree's arrays in place)."""
        self._handle = None
        self._handle_key = None
        self._handle_trees = None

    def _device_model(self):
        """The lists are user-mutable (reference scripts assign model.theta, append stages): re-sync lazily."""
        eng = get_engine()
        key = (eng.device.index, tuple(int(x) for x in self.shape), tuple(id(w) for w in self.classifier),
               tuple(float(t) for t in self.theta))
        if self._handle is None or key != self._handle_key:
            if len(self.classifier) != len(self.theta):
                raise ValueError("classifier and theta must have the same length")
            with eng.torch.cuda.device(eng.device):        # the handle's tables live on the engine's device
                self._handle = ModelHandle(self.shape, self.classifier, self.theta)
            self._handle_key = key
            # the key names the trees by id(): keep them alive as long as the key is, so that a tree that replaces a
            # dropped one (model.classifier.pop(); model.append(DTree(...))) can never be given the same id
            self._handle_trees = list(self.classifier)
        return self._handle

    def _spec(self):
        return resolve_channels(self.channel_opts["channels"])

    def _plan(self, eng, H, W, levels=None, bands=None):
        m, n = int(self.shape[0]), int(self.shape[1])
        assert self.channel_opts["shrink"] in [1, 2], "Shrink factor must be integer 1 <= shrink <= 2"
        return eng.plan(H, W, self.channel_opts, self._spec(), m, n, level_ids=levels, bands=bands)

    # ------------------------------------------------------------------------------------------- reference API
    def channels(self, image):
        """Iterator over the channel pyramid: (chns (u,v,C) float32, scale) (reference model.py:95-103)."""
        yield from _channels.channel_pyramid(image, self.channel_opts)

    def predict_on_image(self, X):
        """All window positions of one channel map (reference model.py:216-259) -> (rs, cs, hs)."""
        X = np.asarray(X)
        u, v, ch_image = X.shape
        m, n, ch_cls = self.shape
        assert ch_image == ch_cls, f"Invalid number of channels. Expected {ch_cls} given {ch_image}."
        hits, stats = get_engine().predict_on_map(self._device_model(), X.astype(np.float32, copy=False))
        self.n_loc += int(stats[0])
        self.n_weak += int(stats[1])
        return hits["r"].astype(np.int64), hits["c"].astype(np.int64), hits["score"].copy()

    def trace_windows(self, X, rs, cs):
        """Leaf index reached in every stage and the float32 score (no rejection) at explicit windows:
        the stage-wise view of DTree.predict_on_image (reference training.py:84-96)."""
        return get_engine().trace(self._device_model(), X, rs, cs)

    def get_boxes(self, r, c, scale) -> Boxes:
        """XYXY boxes in image coordinates (reference model.py:136-147)."""
        r, c = np.asarray(r), np.asarray(c)
        if r.size == 0:
            return Boxes(np.empty((0, 4), "f"))
        m, n = self.shape[:2]
        x1 = c.reshape(-1, 1)
        y1 = r.reshape(-1, 1)
        rects = np.concatenate([x1, y1, x1 + n, y1 + m], axis=1).astype(np.float32)
        return Boxes(rects).normalized(scale=1.0 / scale)

    def _run(self, images, keep_channels=False, levels=None, bands=None):
        """Shared device pipeline: frames [B,H,W] -> (hits, level_counts, plan, chns or None); updates stats.
        `levels`: optional subset of pyramid level indices to compute and scan; `bands`: optional (level, first
        window-tile row, tile rows) triples -- the units of work when one frame is spread over GPUs."""
        eng = get_engine()
        B, H, W = images.shape
        plan = self._plan(eng, H, W, levels, bands)
        handle = self._device_model()
        if plan.n_levels == 0:
            return np.empty(0, dtype=_hit_dtype()), np.zeros((B, 0), np.int32), plan, None
        if keep_channels:
            dev = eng.upload_images(images)
            chns = eng.pyramid(dev, plan)
            hits, counts, stats = eng.cascade(handle, plan, chns, B)
        else:
            chns = None
            hits, counts, stats = eng.run_frames(handle, plan, images)
        self.n_loc += int(stats[:, 0].sum())
        self.n_weak += int(stats[:, 1].sum())
        return hits, counts, plan, chns

    def scan_channels(self, image):
        """Generator of (chns, scale, (r, c, h)) per level (reference model.py:105-134)."""
        _validate_image(image)
        hits, counts, plan, chns = self._run(np.ascontiguousarray(image)[None], keep_channels=True)
        if plan.n_levels == 0:
            return
        host = chns[0].cpu().numpy()
        maps = get_engine().split_levels(host, plan)
        pos = 0
        for lvl in range(plan.n_levels):
            k = int(counts[0, lvl])
            h = hits[pos:pos + k]
            pos += k
            chn = maps[lvl].astype(np.uint8) if self._spec().get("integer") else maps[lvl].copy()
            yield chn, plan.scales[lvl], (h["r"].astype(np.int64), h["c"].astype(np.int64), h["score"].copy())

    @staticmethod
    def _boxes_from_hits(h):
        if h.size == 0:
            b = Boxes(np.empty((0, 4), "f"))
            b.set_field("scores", np.empty(0, np.float32))
            return b
        b = Boxes(np.stack([h["x1"], h["y1"], h["x2"], h["y2"]], axis=1))
        b.set_field("scores", h["score"].copy())
        return b

    def detect(self, image, levels=None) -> Boxes:
        """Detect objects in a 2-D image (reference model.py:149-179) -> Boxes with field "scores", ordered by
        (level, row, column) like the reference's concatenation of per-level results.  `levels` (extension): restrict
        the scan to these pyramid level indices -- the unit of work when one huge frame is spread over several GPUs."""
        _validate_image(image)
        hits, _, _, _ = self._run(np.ascontiguousarray(image)[None], levels=levels)
        return self._boxes_from_hits(hits)

    def predict(self, X):
        """Sample-mode cascade (reference model.py:181-214): X is (K, m, n, C) -> (H, mask); rejected samples get
        H = -inf."""
        X = np.asarray(X)
        assert tuple(X.shape[1:]) == tuple(int(v) for v in self.shape), f"Invalid sample shape {X.shape[1:]}, expected {tuple(self.shape)}"
        return get_engine().predict_samples(self._device_model(), X)

    def detect_batch(self, images, return_hits=False, levels=None, bands=None):
        """Detect on a batch of equally sized frames (array [B,H,W] or list of 2-D arrays) -> list of Boxes.
        With return_hits=True also returns the raw hit records (frame, level, r, c, score, box)."""
        if not isinstance(images, np.ndarray):
            for im in images:
                _validate_image(im)
            images = np.stack(images)
        if images.ndim != 3:
            raise ValueError("detect_batch takes [B,H,W] frames")
        hits, counts, _, _ = self._run(np.ascontiguousarray(images), levels=levels, bands=bands)
        per_frame = counts.sum(axis=1) if counts.size else np.zeros(images.shape[0], np.int64)
        # one (N,4) box array and one score array for the whole batch; the per-frame Boxes are views into them
        all_boxes = np.stack([hits["x1"], hits["y1"], hits["x2"], hits["y2"]], axis=1) if hits.size else np.empty((0, 4), np.float32)
        all_scores = np.ascontiguousarray(hits["score"]) if hits.size else np.empty(0, np.float32)
        out, pos = [], 0
        for b in range(images.shape[0]):
            k = int(per_frame[b])
            boxes = Boxes(all_boxes[pos:pos + k])
            boxes.set_field("scores", all_scores[pos:pos + k])
            out.append(boxes)
            pos += k
        return (out, hits) if return_hits else out

    # ------------------------------------------------------------------------------------------- serialisation
    def as_proto(self, proto):
        """reference model.py:285-306."""
        proto.Clear()
        proto.shape.extend(int(x) for x in self.shape)
        proto.channel_opts.shrink = self.channel_opts["shrink"]
        proto.channel_opts.n_per_oct = self.channel_opts["n_per_oct"]
        proto.channel_opts.smooth = self.channel_opts["smooth"]
        proto.channel_opts.func = symbol_name(self.channel_opts["channels"])
        for weak, theta in self:
            weak.as_proto(proto.classifier.add())
            proto.theta.append(theta)

    @staticmethod
    def from_proto(proto):
        """reference model.py:308-322."""
        shape = tuple(proto.shape)
        channel_opts = {
            "shrink": proto.channel_opts.shrink,
            "n_per_oct": proto.channel_opts.n_per_oct,
            "smooth": proto.channel_opts.smooth,
            "channels": symbol_from_name(proto.channel_opts.func),
        }
        M = Model(shape, channel_opts)
        for weak_proto, theta_proto in zip(proto.classifier, proto.theta):
            M.append(DTree.from_proto(weak_proto), theta_proto)
        return M

    def save(self, filename):
        """zlib-9 compressed proto3 (reference model.py:324-331)."""
        proto = model_pb2.Model()
        self.as_proto(proto)
        with open(filename, "wb") as f:
            f.write(zlib.compress(proto.SerializeToString(), 9))

    @staticmethod
    def load(filename):
        """reference model.py:333-344; unreadable files raise ValueError."""
        with open(filename, "rb") as f:
            data = f.read()
        proto = model_pb2.Model()
        try:
            proto.ParseFromString(zlib.decompress(data))
        except (DecodeError, zlib.error):
            raise ValueError(f"Cannot read model from {filename}")
        return Model.from_proto(proto)


def _hit_dtype():
    from ._native import HIT_DTYPE
    return HIT_DTYPE
