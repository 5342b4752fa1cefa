"""Time the channel pyramid alone (64 device-resident 1080p uint8 frames, config B options); prints ms per launch of
the fused level kernel (CUDA events from the library's profiling hooks).  python profiles/pyramid_time.py [frames]"""
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

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
model = wb.Model.load(os.path.join(ROOT, "tests/golden/configB_model.pb"))
base = [S.synthetic_frame(1000 + i, 1080, 1920) for i in range(min(B, 16))]
frames = np.stack([base[i % len(base)] for i in range(B)])
eng = get_engine()
plan = model._plan(eng, 1080, 1920)
dev = eng.upload_images(frames)
chns = eng.pyramid(dev, plan)
for _ in range(3):
    eng.pyramid(dev, plan, out=chns)
torch.cuda.synchronize()
eng.profile_enable(True)
for _ in range(10):
    eng.pyramid(dev, plan, out=chns)
torch.cuda.synchronize()
ms, n = eng.profile_read()["level_kernel"]
print(json.dumps({"frames": B, "level_kernel_ms": round(ms / n, 4), "checksum": float(chns[0, ::1000].double().sum().item())}))
