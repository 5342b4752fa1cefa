"""Cascade-kernel variant sweep on one GPU (tuning aid, not a bench line).

    python profiles/cascade_sweep.py [frames] [variant-file.json]

Builds the channel pyramid of `frames` synthetic 1080p frames once (config B model), then for every variant (a dict of
WBG_CAS_* environment knobs; the geometry is re-read when the plan and the model handle are created) times the cascade
launch family with CUDA events and checks hits / stats against the first variant.  Prints one JSON line per variant.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C

import numpy as np
import torch

import waldboost_b200 as wb
from waldboost_b200 import _native as N
from waldboost_b200 import synthetic as S
from waldboost_b200.engine import ModelHandle, Plan, get_engine, make_channel_opts

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
DEFAULT = [
    {"name": "default"},
    {"name": "rounds 64/128/128", "WBG_CAS_ROUND_FULL": 64, "WBG_CAS_ROUND_MID": 128},
]
variants = json.load(open(sys.argv[2])) if len(sys.argv) > 2 else DEFAULT
KNOBS = ["WBG_CAS_GEOM384", "WBG_CAS_PACK", "WBG_CAS_ROUND_FULL", "WBG_CAS_ROUND_MID", "WBG_CAS_ROUND_TAIL", "WBG_CAS_ROUND_N1", "WBG_CAS_ROUND_N2",
         "WBG_CAS_ROUND_SOLO"]

model = wb.Model.load(os.path.join(ROOT, os.environ.get("SWEEP_MODEL", "tests/golden/configB_model.pb")))
H, Wd = int(os.environ.get("SWEEP_H", 1080)), int(os.environ.get("SWEEP_W", 1920))
frames = np.stack([S.synthetic_frame(1000 + i, H, Wd) for i in range(B)])
eng = get_engine()
lib = eng.lib
plan0 = model._plan(eng, H, Wd)
dev = eng.upload_images(frames)
chns = eng.pyramid(dev, plan0)
torch.cuda.synchronize()
spec = model._spec() if hasattr(model, "_spec") else None

ref = None
for var in variants:
    for k in KNOBS:
        os.environ.pop(k, None)
    for k, v in var.items():
        if k != "name":
            os.environ[k] = str(v)
    # fresh plan + model handle so the geometry knobs are re-read
    from waldboost_b200.channels import resolve_channels
    sp = resolve_channels(model.channel_opts["channels"])
    plan = Plan(H, Wd, make_channel_opts(model.channel_opts, sp), model.shape[0], model.shape[1])
    mh = ModelHandle(model.shape, model.classifier, model.theta)
    lib.wbg_cascade_counters_enable(1)
    hits, counts, stats = eng.cascade(mh, plan, chns, B)
    cnt = (C.c_uint64 * 16)()
    lib.wbg_cascade_counters_read(cnt)
    lib.wbg_cascade_counters_enable(0)
    cnt = [int(x) for x in cnt]
    # timing: 3 warm-up + 5 timed launches, events on the current stream
    for _ in range(3):
        eng.cascade_launch(mh, plan, chns, B, eng.default_hit_cap(plan, B))
    torch.cuda.synchronize()
    lib.wbg_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for _ in range(reps):
        eng.cascade_launch(mh, plan, chns, B, eng.default_hit_cap(plan, B))
    e1.record()
    torch.cuda.synchronize()
    ms = (C.c_double * 2)()
    ln = (C.c_int64 * 2)()
    lib.wbg_profile_read(ms, ln)
    lib.wbg_profile_enable(0)
    key = (hits.tobytes(), stats.tobytes(), counts.tobytes())
    if ref is None:
        ref = key
    ok = key == ref
    n_weak = int(stats[:, 1].sum())
    execd = sum(cnt[:4])
    out = {"name": var.get("name", "?"), "ok": ok, "kernel_ms_per_frame": ms[1] / max(ln[1], 1) / B,
           "family_ms_per_frame": e0.elapsed_time(e1) / reps / B, "hits": int(hits.size),
           "eval_cost": n_weak / max(int(stats[:, 0].sum()), 1),
           "exec_per_live": execd / max(n_weak, 1), "exec_by_nk": [c / max(n_weak, 1) for c in cnt[:4]],
           "rounds_per_tile": cnt[5] / max(cnt[7], 1), "pool_writes_per_window": cnt[6] / max(int(stats[:, 0].sum()), 1),
           "knobs": {k: v for k, v in var.items() if k != "name"}}
    print(json.dumps(out), flush=True)
    del mh, plan
