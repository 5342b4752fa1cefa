"""Small fixed workload for ncu captures of BASELINE config C: one synthetic 3840x2160 frame, 20x20x10
grad_mag + grad_hist(9) channels, the committed 256-stage model, two passes.  Usage: python profiles/run_once_c.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import waldboost_b200 as wb
from waldboost_b200 import synthetic as S

model = wb.Model.load(os.path.join(ROOT, "tests", "golden", "configC_model.pb"))
frame = S.synthetic_frame(1000, 2160, 3840)
for _ in range(2):
    dt = model.detect(frame)
torch.cuda.synchronize()
print("hits", len(dt), "eval_cost", model.eval_cost)
