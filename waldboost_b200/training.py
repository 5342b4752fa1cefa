"""Inference half of the reference's decision-tree weak classifier (reference waldboost/training.py:23-31,
51-96).  Training (DTree.fit, Learner, rejection schedules) is out of scope of this package."""
import numpy as np


class DTree:
    """Array form of one tree stage: feature u8[N,3]=(r,c,ch), threshold f32[N], left/right i8[N] (-1 at leaves),
    prediction f32[N] (training.py:24-31)."""

    def __init__(self, feature, threshold, left, right, prediction):
        self.feature = np.array([f if f is not None else [0, 0, 0] for f in feature], np.uint8).reshape(-1, 3)
        self.threshold = np.array(threshold, np.float32)
        self.left = np.array(left, np.int8)
        self.right = np.array(right, np.int8)
        self.prediction = np.array(prediction, np.float32)
        self.node = self.left >= 0
        self.node_idx = np.flatnonzero(self.node)

    @staticmethod
    def from_proto(proto):
        """training.py:51-59."""
        ftr = np.array(proto.feature).reshape((-1, 3))
        ftr = [tuple(x) if x[0] >= 0 else None for x in ftr]
        return DTree(ftr, np.array(proto.threshold), np.array(proto.left), np.array(proto.right),
                     np.array(proto.prediction))

    def as_proto(self, proto):
        """training.py:60-72 (leaf features are written as 0,0,0: rows of a uint8 array are never None)."""
        proto.Clear()
        proto.feature.extend(int(x) for x in self.feature.reshape(-1))
        proto.threshold.extend(float(x) for x in self.threshold)
        proto.left.extend(int(x) for x in self.left)
        proto.right.extend(int(x) for x in self.right)
        proto.prediction.extend(float(x) for x in self.prediction)

    def predict_on_image(self, X, rs, cs):
        """training.py:84-96 on the GPU: prediction of this single tree at windows (rs, cs) of channel map X."""
        from .model import Model
        m = int(self.feature[:, 0].max()) + 1
        n = int(self.feature[:, 1].max()) + 1
        X = np.ascontiguousarray(X, np.float32)
        one = Model((m, n, X.shape[2]), None)
        one.append(self, -np.inf)
        _, score = one.trace_windows(X, rs, cs)
        return score
