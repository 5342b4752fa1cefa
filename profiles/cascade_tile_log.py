"""Per-tile lifetime and work of the cascade kernel (debug log) on one synthetic 1080p frame per launch.
python profiles/cascade_tile_log.py [frames]"""
import ctypes as C
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

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
model = wb.Model.load(os.path.join(ROOT, "tests/golden/configB_model.pb"))
frames = np.stack([S.synthetic_frame(1000 + i, 1080, 1920) for i in range(B)])
eng = get_engine()
lib = eng.lib
lib.wbg_cascade_tile_log.restype = C.c_int
lib.wbg_cascade_tile_log.argtypes = [C.c_void_p, C.c_int64]
plan = model._plan(eng, 1080, 1920)
chns = eng.pyramid(eng.upload_images(frames), plan)
mh = model._device_model()
for _ in range(2):
    eng.cascade(mh, plan, chns, B)
lib.wbg_cascade_counters_enable(1)
hits, counts, stats = eng.cascade(mh, plan, chns, B)
torch.cuda.synchronize()
nt = min(16384, 1670 * B)
buf = np.zeros(2 * nt, np.uint64)
lib.wbg_cascade_tile_log(buf.ctypes.data, nt)
lib.wbg_cascade_counters_enable(0)
life = buf[0::2].astype(np.float64)
work = buf[1::2].astype(np.float64)
ok = life > 0
life, work = life[ok], work[ok]
print(json.dumps({"tiles": int(ok.sum()), "life_kcycles": {q: round(float(np.percentile(life, q)) / 1e3, 1) for q in (10, 50, 90, 99, 100)},
                  "work_kws": {q: round(float(np.percentile(work, q)) / 1e3, 1) for q in (10, 50, 90, 99, 100)},
                  "life_sum_Mcycles": round(life.sum() / 1e6, 1), "work_sum_Mws": round(work.sum() / 1e6, 1),
                  "corr": round(float(np.corrcoef(life, work)[0, 1]), 3)}))
# lifetime share by work decile
order = np.argsort(work)
for lo in range(0, 100, 10):
    sel = order[int(len(order) * lo / 100):int(len(order) * (lo + 10) / 100)]
    print(json.dumps({"work_decile": lo // 10, "work_kws_mean": round(float(work[sel].mean()) / 1e3, 1), "life_kcycles_mean": round(float(life[sel].mean()) / 1e3, 1),
                      "life_share": round(float(life[sel].sum() / life.sum()), 3), "work_share": round(float(work[sel].sum() / work.sum()), 3),
                      "cycles_per_ws": round(float(life[sel].sum() / work[sel].sum()), 3)}))
