"""Small fixed workload for compute-sanitizer (memcheck / racecheck / initcheck, one tool per run): the three stage
encodings of the cascade kernel (canonical depth-2 stages from the constant bank, complete depth-4 records staged in
shared memory, node records of arbitrary topology), the 4-bin uint8 and the generic channel kernels, hit emission,
multi-model detect (constant-bank hand-over between two models) and the explicit-window entry points.
python profiles/sanitize_workload.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np

import waldboost_b200 as wb
from waldboost_b200 import channels as CH
from waldboost_b200 import synthetic as S

G = os.path.join(ROOT, "tests", "golden")
frame = S.synthetic_frame(1000, 96, 128)
small = wb.Model.load(os.path.join(G, "small_model.pb"))            # depth-2 (constant bank)
generic = wb.Model.load(os.path.join(G, "generic_model.pb"))        # depth 1 / 3 / unbalanced
A, B = wb.Model.load(os.path.join(G, "multi_A_model.pb")), wb.Model.load(os.path.join(G, "multi_B_model.pb"))
n = 0
n += len(small.detect(frame))
n += len(generic.detect(frame))
n += len(wb.detect(S.synthetic_frame(1000, 200, 260), A, B, response_scale=[1.0, 0.5]))
# depth-4 (DK4) and a node-record model with 5 levels; float32 frame and grad_mag channels through the generic kernels
from helpers import make_model
opts = dict(shrink=2, n_per_oct=4, smooth=1, channels=CH.grad_hist)
d4 = make_model((12, 12, 4), opts, 12, 4, frame, keep_total=5e-2)
d5 = make_model((12, 12, 4), opts, 6, 5, frame, keep_total=0.2)
n += len(d4.detect(frame)) + len(d5.detect(frame))
mag = make_model((10, 10, 1), dict(shrink=2, n_per_oct=2, smooth=1, channels=CH.grad_mag), 6, 2, frame, keep_total=0.2)
n += len(mag.detect(frame.astype(np.float32)))
out = small.detect_batch(np.stack([frame] * 10))                     # pipelined chunks on two streams
X = next(iter(small.channels(frame)))[0]
rs, cs = np.array([0, 3, 7]), np.array([1, 2, 9])
leaf, score = small.trace_windows(X, rs, cs)
print("sanitize workload done:", n, len(out), leaf.shape)
