"""waldboost_b200 -- B200-native (sm_100a CUDA) implementation of the WaldBoost detector's inference hot path,
with the package-level API of the reference (reference waldboost/__init__.py:43-130):

    import waldboost_b200 as wb
    model = wb.load("model.pb")
    boxes = model.detect(image)                 # or wb.detect(image, model_a, model_b)

Training, sample pools, the FPGA variant and evaluation helpers of the reference are out of scope.
The CUDA library (libwbg.so, include/wbg.h) is required: there is no CPU fallback.
"""
import numpy as np

from . import channels
from . import fpga
from .boxes import Boxes, concatenate
from .model import Model
from .training import DTree

__version__ = "0.2.0+b200.r1"

load = load_model = Model.load


def save_model(model: Model, filename):
    """Save model to file.  See Model.save (reference __init__.py:67-72)."""
    model.save(filename)


save = save_model


def detect(image: np.ndarray, *models: Model, channel_opts: dict = None, response_scale=None) -> Boxes:
    """Detect with several models sharing one channel pyramid (reference __init__.py:75-130).

    Returns Boxes with fields "scores" (model response times response_scale[k]) and "label" (index of the model),
    ordered by (level, model, row, column) like the reference's nested loops.  The reference's `np.int` at
    __init__.py:128 no longer exists in NumPy >= 1.24; labels are int64 here.
    """
    from .channels import _validate_image, resolve_channels
    from .engine import get_engine
    if not models:
        raise ValueError("detect() needs at least one model")
    channel_opts = channel_opts or models[0].channel_opts
    if response_scale is None:
        response_scale = [1] * len(models)
    response_scale = np.array(response_scale, "f")
    if response_scale.size != len(models):
        raise ValueError("Wrong response_scale parameter")
    _validate_image(image)
    assert channel_opts["shrink"] in [1, 2], "Shrink factor must be integer 1 <= shrink <= 2"
    eng = get_engine()
    spec = resolve_channels(channel_opts["channels"])
    image = np.ascontiguousarray(image)
    H, W = image.shape
    dev = eng.upload_images(image[None])
    chns, per_model, n_levels = None, [], 0
    for model in models:
        m, n, ch_cls = model.shape
        plan = eng.plan(H, W, channel_opts, spec, int(m), int(n))
        assert plan.C == ch_cls, f"Invalid number of channels. Expected {ch_cls} given {plan.C}."
        n_levels = plan.n_levels
        if n_levels == 0:
            break
        if chns is None:
            chns = eng.pyramid(dev, plan)          # one pyramid for all models (the layout does not depend on m, n)
        hits, counts, stats = eng.cascade(model._device_model(), plan, chns, 1)
        model.n_loc += int(stats[0, 0])
        model.n_weak += int(stats[0, 1])
        per_model.append((hits, np.concatenate([[0], np.cumsum(counts[0])])))
    dt_boxes = []
    for lvl in range(n_levels):
        for k, (hits, starts) in enumerate(per_model):
            h = hits[starts[lvl]:starts[lvl + 1]]
            if h.size == 0:
                continue
            boxes = Boxes(np.stack([h["x1"], h["y1"], h["x2"], h["y2"]], axis=1))
            boxes.set_field("scores", h["score"] * response_scale[k])
            boxes.set_field("label", np.full(h.size, k, dtype=np.int64))
            dt_boxes.append(boxes)
    if not dt_boxes:
        out = Boxes(np.empty((0, 4), "f"))
        out.set_field("scores", np.empty(0, np.float32))
        out.set_field("label", np.empty(0, np.int64))
        return out
    return concatenate(dt_boxes, ["scores", "label"])
