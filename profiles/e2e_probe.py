import os, sys, time, json
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import waldboost_b200 as wb
from waldboost_b200 import synthetic as S
B=64
model = wb.Model.load("/root/repo/tests/golden/configB_model.pb")
pinned = torch.empty((B,1080,1920), dtype=torch.uint8, pin_memory=True)
frames = pinned.numpy()
for i in range(B): frames[i] = S.synthetic_frame(1000+i, 1080, 1920)
for chunk, grow in ((8, 1), (8, 2), (8, 4), (4, 4), (4, 8), (16, 2), (8, 1)):
    os.environ["WBG_PIPE_CHUNK"] = str(chunk)
    os.environ["WBG_PIPE_GROW"] = str(grow)
    for _ in range(2): model.detect_batch(frames)
    torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(5): out, hits = model.detect_batch(frames, return_hits=True)
    torch.cuda.synchronize()
    dt=(time.perf_counter()-t0)/5
    print(json.dumps({"chunk": chunk, "grow": grow, "ms_per_step": round(dt*1e3,2), "fps": round(B/dt), "hits": int(hits.size)}), flush=True)
import cProfile, pstats
os.environ["WBG_PIPE_CHUNK"]="8"; os.environ["WBG_PIPE_GROW"]="4"
pr=cProfile.Profile(); pr.enable()
for _ in range(3): model.detect_batch(frames, return_hits=True)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
