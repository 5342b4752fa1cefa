"""TEST INFRASTRUCTURE ONLY -- the CPU arm of bench.py (`cpu_baseline` and `--impl reference`).

Times the reference's own detect() path on the host cores WITHOUT importing the product package (no waldboost_b200,
no libwbg.so in the process):
  * when the unmodified reference is importable (/root/reference, i.e. in the build container) each worker process runs
    the reference's `Model.detect` on whole frames under the shims of oracle/ref_harness.py -> kind "reference";
  * elsewhere (the GPU box has no /root/reference) each worker runs the NumPy restatement oracle/wb_oracle.py -> kind
    "port"; a frame's pyramid levels are then additionally split over a few workers so that a step stays short
    (levels are independent of one another, reference channels.py:125-132).
The model file is parsed here with protobuf descriptors built from the reference's waldboost/model.proto:1-23
(ref_harness._build_model_pb2) and zlib (model.py:324-344); frames follow the recipe of SURVEY.md 8d, restated in
`synthetic_frame` and pinned against the product's generator by tests/test_bench_contract.py.
"""
import os
import sys
import time
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

CPU_GROUP = 4      # port only: worker processes sharing one frame (its levels are split between them; largest level = 16 %)


# ----------------------------------------------------------------------------------------------- inputs
def synthetic_frame(seed, H, W):
    """SURVEY.md 8d after reference waldboost/utils.py:81-97: bright axis-aligned squares on a dark canvas plus uniform
    noise, clipped and cast to uint8 (bit-identical to waldboost_b200.synthetic.synthetic_frame)."""
    rng = np.random.default_rng(seed)
    canvas = np.zeros((H, W), np.float64)
    k = int(rng.integers(0, 1 + int(H * W / 65536 * 2)))
    hi = max(60 * H / 256, 31)
    for _ in range(k):
        side = int(rng.uniform(30, hi))
        y = int(rng.integers(0, max(H - side, 1)))
        x = int(rng.integers(0, max(W - side, 1)))
        canvas[y:y + side, x:x + side] += rng.uniform(0.2, 1.0)
    canvas += rng.random((H, W)) * 0.3 * rng.random()
    return (np.clip(canvas, 0, 1) * 255).astype(np.uint8)


def read_model_pb(path):
    """.pb written by Model.save (model.py:285-344): zlib + proto3 -> dict(shape, shrink, n_per_oct, smooth, func,
    trees [(feature [N,3] u8, threshold, left, right, prediction)], theta)."""
    import ref_harness
    pb2 = ref_harness._build_model_pb2()
    proto = pb2.Model()
    with open(path, "rb") as f:
        proto.ParseFromString(zlib.decompress(f.read()))
    trees = []
    for w in proto.classifier:
        trees.append((np.array(w.feature, np.uint8).reshape(-1, 3), np.array(w.threshold, np.float32),
                      np.array(w.left, np.int8), np.array(w.right, np.int8), np.array(w.prediction, np.float32)))
    return dict(shape=tuple(proto.shape), shrink=proto.channel_opts.shrink, n_per_oct=proto.channel_opts.n_per_oct,
                smooth=proto.channel_opts.smooth, func=proto.channel_opts.func, trees=trees,
                theta=[float(np.float32(t)) for t in proto.theta])


def oracle_cascade(desc, profile="wald"):
    import wb_oracle as O
    fn = getattr(O, desc["func"].rsplit(".", 1)[-1])
    opts = dict(shrink=desc["shrink"], n_per_oct=desc["n_per_oct"], smooth=desc["smooth"], channels=fn)
    Cs = O.Cascade(desc["shape"], opts)
    for (f, t, l, r, p), th in zip(desc["trees"], desc["theta"]):
        Cs.append(O.DTree([tuple(x) for x in f], t, l, r, p), -np.inf if profile == "dense" else th)
    return Cs


def reference_model(desc, profile="wald"):
    """the same cascade as an UNMODIFIED reference Model (needs /root/reference)."""
    import ref_harness
    ref_harness.import_reference()
    from waldboost import channels as rch
    from waldboost.model import Model as RModel
    from waldboost.training import DTree as RDTree
    fn = getattr(rch, desc["func"].rsplit(".", 1)[-1])
    M = RModel(desc["shape"], dict(shrink=desc["shrink"], n_per_oct=desc["n_per_oct"], smooth=desc["smooth"], channels=fn))
    for (f, t, l, r, p), th in zip(desc["trees"], desc["theta"]):
        M.append(RDTree([tuple(x) for x in f], t, l, r, p), -np.inf if profile == "dense" else th)
    return M


def level_shards(desc, h, w, groups):
    """greedy longest-processing-time split of the pyramid levels (cost = channel pixels) over `groups` workers."""
    import wb_oracle as O
    costs = []
    for oh, ow in [x.shape for x in O.image_octaves(np.zeros((h, w), np.uint8))]:
        for i in range(desc["n_per_oct"]):
            nh, nw = O.level_size(oh, ow, i, desc["n_per_oct"], desc["shrink"])
            costs.append((nh // desc["shrink"]) * (nw // desc["shrink"]))
    order = sorted(range(len(costs)), key=lambda l: (-costs[l], l))
    load, out = [0.0] * groups, [[] for _ in range(groups)]
    for l in order:
        r = min(range(groups), key=lambda k: (load[k], k))
        out[r].append(l)
        load[r] += costs[l]
    return [sorted(x) for x in out]


def kind(force=None):
    """"reference" when the unmodified reference can be imported, else "port"; `force` = "port" keeps to the restatement."""
    import ref_harness
    if force == "port":
        return "port"
    ok = ref_harness.reference_available()
    if force == "reference" and not ok:
        raise RuntimeError("the reference is not present (looked under %s)" % ref_harness.REFERENCE_ROOT)
    return "reference" if ok else "port"


# ----------------------------------------------------------------------------------------------- workers
def _worker(job):
    """one worker process: detect() on (some levels of) one frame; returns (seconds, hits, n_loc, n_weak)."""
    model_path, seed, levels, profile, h, w, which, crop = job
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("NUMBA_NUM_THREADS", "1")
    desc = read_model_pb(model_path)
    frame = synthetic_frame(seed, h, w)
    if crop:
        frame = np.ascontiguousarray(frame[crop[0]:crop[1], crop[2]:crop[3]])
    if which == "reference":
        M = reference_model(desc, profile)
        M.detect(synthetic_frame(1, 64, 96))           # untimed: Numba JIT of the reference's kernels (9-15 s)
        M.reset()
        t0 = time.perf_counter()
        n = len(M.detect(frame))
        return time.perf_counter() - t0, n, M.n_loc, M.n_weak
    Cs = oracle_cascade(desc, profile)
    t0 = time.perf_counter()
    n = Cs.detect(frame, levels)[1].size
    return time.perf_counter() - t0, n, Cs.n_loc, Cs.n_weak


class CpuPool:
    """Worker processes over frames -- the reference's own parallel pattern (multiprocessing.Pool over files,
    scripts/waldboost-detect.py:65).  Spawned, so no CUDA context is inherited."""

    def __init__(self, workers, model_path, force_kind=None):
        import multiprocessing as mp
        self.kind = kind(force_kind)
        self.model_path = model_path
        self.desc = read_model_pb(model_path)
        self.group = 1 if self.kind == "reference" else min(CPU_GROUP, workers)
        self.frames = max(1, workers // self.group)
        self.workers = self.frames * self.group
        self.pool = mp.get_context("spawn").Pool(self.workers)

    def step(self, profile, h, w, seed0=1000):
        shards = level_shards(self.desc, h, w, self.group) if self.group > 1 else [None]
        jobs = [(self.model_path, seed0 + f, shards[g], profile, h, w, self.kind, None)
                for f in range(self.frames) for g in range(self.group)]
        res = self.pool.map(_worker, jobs, chunksize=1)
        busy = max(r[0] for r in res)            # slowest worker's detect time (excludes start-up, JIT and frame synthesis)
        return self.frames, busy, sum(r[1] for r in res), sum(r[2] for r in res), sum(r[3] for r in res)

    def single_thread(self, profile, h, w, crop, seed=1000):
        """one worker, one core: detect() on a crop of one frame -> (frames/s scaled by the pixel share, description)."""
        t, n, n_loc, n_weak = self.pool.apply(_worker, ((self.model_path, seed, None, profile, h, w, self.kind, crop),))
        share = (crop[1] - crop[0]) * (crop[3] - crop[2]) / float(h * w)
        return share / t, f"{crop[3] - crop[2]}x{crop[1] - crop[0]} crop of one frame on 1 core in {t:.1f} s ({n_weak / max(n_loc, 1):.1f} stages per window), scaled by its {share:.2f} pixel share"

    def sample(self, h, w):
        what = ("the unmodified reference's Model.detect (waldboost/model.py:149-179, imported under oracle/ref_harness.py)"
                if self.kind == "reference" else "oracle/wb_oracle.py Cascade.detect (NumPy restatement of the reference)")
        split = "" if self.group == 1 else f" (each frame's pyramid levels split over {self.group} workers)"
        return f"{self.frames} frames of {w}x{h} per step on {self.workers} worker processes{split}, {what}"

    def close(self):
        self.pool.close()
        self.pool.join()
