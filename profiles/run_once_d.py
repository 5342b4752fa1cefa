"""Small fixed workload for ncu captures of the depth-4 cascade path (config D shape): 32 synthetic 640x480 frames, 2048-stage
depth-4 model calibrated like profiles/run_configs.py."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "profiles"))
import numpy as np
import torch

import waldboost_b200 as wb
from waldboost_b200 import synthetic as S
from run_configs import calibrated_model

opts = dict(shrink=2, n_per_oct=8, smooth=1, channels=wb.channels.grad_hist)
frames = S.synthetic_frames(32, 480, 640)
M = calibrated_model((12, 12, 4), opts, 2048, 4, frames[0], 1e-4)
for _ in range(2):
    M.reset()
    out, hits = M.detect_batch(frames, return_hits=True)
torch.cuda.synchronize()
print("hits", hits.size, "eval_cost", M.eval_cost)
