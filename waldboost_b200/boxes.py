"""Minimal stand-in for the third-party `bbx.Boxes` container the reference returns from detect()
(reference waldboost/model.py:15-16,139-147,177-179; `bbx` itself is not part of the reference tree).

Only the surface the detect() path touches is provided: construction from an (N,4) XYXY array, named
per-box fields, `normalized(shift, scale)`, `get()`, indexing and `concatenate`.
"""
import numpy as np


class Boxes:
    def __init__(self, C, **fields):
        self.C = np.asarray(C, np.float32).reshape(-1, 4)
        self.fields = {}
        for k, v in fields.items():
            self.set_field(k, v)

    def get(self):
        """(N,4) float32 array [xmin, ymin, xmax, ymax]."""
        return self.C

    def normalized(self, shift=(0, 0), scale=1):
        """Coordinates shifted then multiplied by `scale` in float32 (used as scale=1/level_scale, model.py:147)."""
        B = Boxes((self.C - np.tile(np.asarray(shift, np.float32), 2)) * np.float32(scale))
        B.fields = dict(self.fields)
        return B

    def set_field(self, name, value):
        value = np.asarray(value)
        if value.shape[:1] != (len(self),):
            raise ValueError(f"field {name!r} must have {len(self)} rows, got shape {value.shape}")
        self.fields[name] = value

    def get_field(self, name):
        return self.fields[name]

    def has_field(self, name):
        return name in self.fields

    def get_fields(self):
        return list(self.fields)

    def __len__(self):
        return self.C.shape[0]

    def __bool__(self):
        return len(self) > 0

    def __getitem__(self, i):
        idx = np.atleast_1d(np.arange(len(self))[i])
        B = Boxes(self.C[idx])
        B.fields = {k: v[idx] for k, v in self.fields.items()}
        return B

    def __repr__(self):
        return f"Boxes(n={len(self)}, fields={list(self.fields)})"


def concatenate(boxes, fields=None):
    """bbx.concatenate (model.py:179, __init__.py:130): stack boxes and the named fields in order."""
    boxes = list(boxes)
    if not boxes:
        return Boxes(np.empty((0, 4), np.float32))
    if fields is None:
        fields = boxes[0].get_fields()
    out = Boxes(np.concatenate([b.get() for b in boxes], axis=0))
    for f in fields:
        out.set_field(f, np.concatenate([np.atleast_1d(b.get_field(f)) for b in boxes]))
    return out
