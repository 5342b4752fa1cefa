"""Kernel times of BASELINE config C (one 3840x2160 frame, committed 256-stage 20x20x10 model) from the library's
CUDA-event hooks: level kernel and cascade kernel, ms per frame.  Usage: python profiles/config_c_kernels.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import waldboost_b200 as wb
from waldboost_b200 import synthetic as S
from waldboost_b200.engine import get_engine

model = wb.Model.load(os.path.join(ROOT, "tests", "golden", "configC_model.pb"))
frame = S.synthetic_frame(1000, 2160, 3840)
eng = get_engine()
for _ in range(3):
    model.detect(frame)
torch.cuda.synchronize()
eng.profile_enable(True)
reps = 10
for _ in range(reps):
    dt = model.detect(frame)
torch.cuda.synchronize()
prof = eng.profile_read()
eng.profile_enable(False)
print(json.dumps({"hits": len(dt), "level_kernel_ms": prof["level_kernel"][0] / prof["level_kernel"][1],
                  "cascade_kernel_ms": prof["cascade_kernel"][0] / prof["cascade_kernel"][1]}))
