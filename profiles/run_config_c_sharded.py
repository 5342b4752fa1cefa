"""BASELINE config C on N GPUs: ONE 3840x2160 frame, 20x20x10 grad_mag+grad_hist model, pyramid levels sharded across the
ranks (greedy by channel pixels), host-side gather of the hit records -- no data-path collective.  Rank 0 checks the
gathered result against its own single-GPU full detect and prints one JSON line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/run_config_c_sharded.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "profiles"))
import numpy as np
import torch
import torch.distributed as dist

import waldboost_b200 as wb
from waldboost_b200 import sharding
from waldboost_b200 import synthetic as S
from run_configs import calibrated_model

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
opts = dict(shrink=2, n_per_oct=8, smooth=1, channels=wb.channels.grad_mag_hist)
frame = S.synthetic_frame(1000, 2160, 3840)
M = calibrated_model((20, 20, 10), opts, 256, 2, S.synthetic_frame(1000, 540, 960), 1e-4)     # same seeds on every rank
from waldboost_b200.engine import get_engine, plan_geometry
plan = plan_geometry(2160, 3840, opts, M._spec(), 20, 20)
costs = [lv.u * lv.v for lv in plan.levels]


def detect_levels(ids):
    M.reset()
    _, hits = M.detect_batch(frame[None], return_hits=True, levels=ids)
    return hits, (M.n_loc, M.n_weak)


sharding.detect_level_sharded(detect_levels, costs)               # warm-up (plans, buffers, the group's object gather)
torch.cuda.synchronize()
t_sharded = []
for _ in range(5):
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    hits, stats = sharding.detect_level_sharded(detect_levels, costs)
    torch.cuda.synchronize()
    t_sharded.append(time.perf_counter() - t0)
t_sharded = float(np.median(t_sharded))
if rank == 0:
    detect_levels(list(range(len(costs))))
    t0 = time.perf_counter()
    full, full_stats = detect_levels(list(range(len(costs))))
    t_full = time.perf_counter() - t0
    loads = [sum(costs[l] for l in p) for p in sharding.assign_levels(costs, world)]
    print(json.dumps({"config": "C: one 3840x2160 frame, 20x20x10, levels sharded over ranks", "n_gpus": world,
                      "identical_to_single_gpu": bool(np.array_equal(hits, full) and tuple(stats) == tuple(full_stats)),
                      "hits": int(hits.size), "ms_sharded_incl_gather": t_sharded * 1e3, "ms_single_gpu": t_full * 1e3,
                      "max_rank_share_of_pixels": max(loads) / sum(costs)}), flush=True)
if world > 1:
    dist.destroy_process_group()
