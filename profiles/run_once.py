"""Small fixed workload for ncu captures: 4 synthetic 1080p frames through the channel pyramid and the cascade
(config B model), twice.  Usage: python profiles/run_once.py [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import waldboost_b200 as wb
from waldboost_b200 import synthetic as S
from waldboost_b200.engine import get_engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
model = wb.Model.load(os.path.join(ROOT, "tests", "golden", "configB_model.pb"))
frames = np.stack([S.synthetic_frame(1000 + i, 1080, 1920) for i in range(B)])
eng = get_engine()
plan = model._plan(eng, 1080, 1920)
dev = eng.upload_images(frames)
for _ in range(2):
    chns = eng.pyramid(dev, plan)
    hits, counts, stats = eng.cascade(model._device_model(), plan, chns, B)
torch.cuda.synchronize()
print("hits", hits.size, "eval_cost", stats[:, 1].sum() / stats[:, 0].sum())
