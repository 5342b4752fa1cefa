"""Where does the cascade kernel's time go?  Times the kernel on the config B model truncated to its first T' stages
(the last kept stage gets theta = +inf, so nothing survives and no hit is emitted) for growing T'; the differences are
the cost of each block of stages with every tile of the batch in flight.  python profiles/cascade_phase_cost.py [frames]"""
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
from waldboost_b200.channels import resolve_channels
from waldboost_b200.engine import ModelHandle, Plan, get_engine, make_channel_opts

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
model = wb.Model.load(os.path.join(ROOT, "tests/golden/configB_model.pb"))
frames = np.stack([S.synthetic_frame(1000 + i, 1080, 1920) for i in range(B)])
eng = get_engine()
lib = eng.lib
plan0 = model._plan(eng, 1080, 1920)
chns = eng.pyramid(eng.upload_images(frames), plan0)
torch.cuda.synchronize()
sp = resolve_channels(model.channel_opts["channels"])
plan = Plan(1080, 1920, make_channel_opts(model.channel_opts, sp), model.shape[0], model.shape[1])
prev = None
for T in [1, 16, 32, 64, 96, 128, 192, 256, 384, 512, 768, 1024]:
    thetas = list(model.theta[:T])
    thetas[-1] = float("inf")
    mh = ModelHandle(model.shape, model.classifier[:T], thetas)
    lib.wbg_cascade_counters_enable(1)
    hits, counts, stats = eng.cascade(mh, plan, chns, B)
    cnt = (C.c_uint64 * 16)()
    lib.wbg_cascade_counters_read(cnt)
    lib.wbg_cascade_counters_enable(0)
    cnt = [int(x) for x in cnt]
    for _ in range(3):
        eng.cascade_launch(mh, plan, chns, B, 1024)
    torch.cuda.synchronize()
    lib.wbg_profile_enable(1)
    for _ in range(5):
        eng.cascade_launch(mh, plan, chns, B, 1024)
    torch.cuda.synchronize()
    ms = (C.c_double * 2)()
    ln = (C.c_int64 * 2)()
    lib.wbg_profile_read(ms, ln)
    lib.wbg_profile_enable(0)
    us = 1e3 * ms[1] / max(ln[1], 1) / B
    n_weak = int(stats[:, 1].sum()) / B
    execd = sum(cnt[:4]) / B
    row = {"T": T, "us_per_frame": round(us, 2), "live_Mws_per_frame": round(n_weak / 1e6, 2), "exec_Mws_per_frame": round(execd / 1e6, 2),
           "hits": int(hits.size)}
    if prev:
        dus, dlive, dexec = us - prev[0], n_weak - prev[1], execd - prev[2]
        row.update({"d_us": round(dus, 2), "d_live_Mws": round(dlive / 1e6, 2), "d_exec_Mws": round(dexec / 1e6, 2),
                    "ns_per_live_kws": round(1e3 * dus / max(dlive / 1e3, 1e-9), 3),
                    "sm_cycles_per_exec_warp_stage": round(dus * 1e-6 * 1.965e9 * 148 / max(dexec / 32, 1e-9), 2)})
    print(json.dumps(row), flush=True)
    prev = (us, n_weak, execd)
    del mh
