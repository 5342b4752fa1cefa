"""Does keeping the channel block in L2 between the pyramid and the cascade help?  
Runs pyramid+cascade over 64 device-resident 1080p frames in sub-batches of 64 / 16 / 8 / 4 / 2 / 1 frames issued back to
back on one stream (a sub-batch of 2 frames is 104 MB of channels, the L2 holds 126 MB) and reports ms per 64 frames.
python profiles/l2_residency.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import waldboost_b200 as wb
from waldboost_b200 import synthetic as S
from waldboost_b200.engine import get_engine

B = 64
model = wb.Model.load(os.path.join(ROOT, "tests/golden/configB_model.pb"))
frames = np.stack([S.synthetic_frame(1000 + i, 1080, 1920) for i in range(B)])
eng = get_engine()
plan = model._plan(eng, 1080, 1920)
handle = model._device_model()
dev = eng.upload_images(frames)
chns = eng.pyramid(dev, plan)
for sub in (64, 16, 8, 4, 2, 1):
    cap = eng.default_hit_cap(plan, sub)

    def step():
        for lo in range(0, B, sub):
            c = chns[lo:lo + sub]
            eng.pyramid(dev[lo:lo + sub], plan, out=c)
            eng.cascade_launch(handle, plan, c, sub, cap)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        step()
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"frames_per_sub_batch": sub, "channel_MB_per_sub_batch": round(sub * plan.chn_floats * 4 / 1e6, 1),
                      "ms_per_64_frames": round(e0.elapsed_time(e1) / 5, 3)}), flush=True)
