#!/usr/bin/env python
"""bench.py -- 1080p frames/s of Model.detect (BASELINE.json metric) on B200, with the HBM roofline of the two dominant
kernels and the CPU reference path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch 64] [--profile wald|dense] [--impl reference]

Workload (BASELINE.json configs[1] / configs[4]): a batch of 64 synthetic 1920x1080 uint8 frames per GPU and the
12x12x4 grad_hist model (shrink 2, n_per_oct 8, smooth 1) with 1024 depth-2 stages and 'wald' rejection thresholds
(tests/golden/configB_model.pb, calibrated with the reference).  One "step" = detect() over one batch on
every rank.  N > 1: one process per GPU (torchrun), frames sharded by image, no data-path collective; the process group
is used only for the barrier and the max-over-ranks of the step time ("scaling": "weak").

  value     frames/s with the frames already resident in HBM: fused channel pyramid + cascade + hit emission.
  e2e       frames/s through the public API (Model.detect_batch on pinned HOST frames): H2D copy of the frames and D2H
            read of the hits inside the timed region.
  roofline  for the kernel with the larger share of the step: ALGORITHMIC bytes per launch / its CUDA-event time;
            `roofline_pyramid` and `roofline_cascade` give both kernels.
  cpu_baseline  the reference's CPU path on the host cores (oracle/cpu_arm.py): the unmodified reference where
            /root/reference exists ("kind": "reference"), its NumPy restatement oracle/wb_oracle.py elsewhere ("port"),
            image-sharded over worker processes on a bounded sample of the same workload, plus a single-core figure.

--impl reference times only that CPU path and prints the same line with "impl": "reference".
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODEL_B = os.path.join(ROOT, "tests", "golden", "configB_model.pb")
H, W = 1080, 1920
METRIC = "1080p frames/sec for Model.detect"
UNIT = "frames/s"


# ----------------------------------------------------------------------------------------------- CPU reference path
# Everything the CPU arm runs lives under oracle/ (oracle/cpu_arm.py): the unmodified reference when /root/reference is
# present (build container), its NumPy restatement otherwise (GPU box).  Neither this arm nor its worker processes import
# waldboost_b200 or map libwbg.so.
def _cpu_arm():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_arm   # (checker / CPU baseline only)
    return cpu_arm


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


SINGLE_CROP = (270, 810, 480, 1440)     # centre 540x960 of a 1080p frame: the bounded single-thread sample


def cpu_measure(args, steps, warm):
    """N-process and single-thread frames/s of the reference's CPU path on a bounded sample of the workload."""
    A = _cpu_arm()
    cores = max(1, min(host_cores(), args.cpu_workers or 64))
    pool = A.CpuPool(cores, MODEL_B, None if args.cpu_kind == "auto" else args.cpu_kind)
    for _ in range(warm):                        # untimed: a small frame per worker (imports, page-in, JIT if any)
        pool.step(args.profile, 135, 240)
    t, frames, n_loc, n_weak = 0.0, 0, 0, 0
    for _ in range(steps):
        n, busy, _, nl, nw = pool.step(args.profile, H, W)
        t += busy; frames += n; n_loc += nl; n_weak += nw
    single, single_txt = (None, None)
    if not args.no_single_thread:
        single, single_txt = pool.single_thread(args.profile, H, W, SINGLE_CROP)
    sample = pool.sample(H, W)
    kind, workers = pool.kind, pool.workers
    pool.close()
    fps = frames / t
    return {"value": fps, "unit": UNIT, "cores": workers, "kind": kind,
            "sample": f"{sample}, {steps} step(s), {n_weak / max(n_loc, 1):.1f} stages per window, {t:.1f} s",
            "single_thread": {"value": single, "unit": UNIT, "cores": 1, "sample": single_txt}}, t


def run_reference(args):
    """--impl reference: the reference's CPU path on all host cores.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cpu, t = cpu_measure(args, args.steps, max(args.warmup, 1))
    fps = cpu["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"Model.detect, 1920x1080 uint8 frames, 12x12x4 grad_hist model, 1024 depth-2 stages, {args.profile} thetas "
                               "(BASELINE configs[1]); each step a bounded sample: one frame per worker group",
                   "profile": args.profile},
        "cpu_baseline": cpu,
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        # evidence that this arm is the CPU path only: shared objects of the product mapped into this process
        "native_so_loaded": _mapped_product_libraries(),
    }
    print(json.dumps(line), flush=True)


def _mapped_product_libraries():
    try:
        with open("/proc/self/maps") as f:
            return sorted({l.split()[-1] for l in f if "libwbg" in l})
    except OSError:
        return None


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU with NVML while the timed region runs."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------- GPU arm
def algorithmic_bytes(plan, dtype_bytes, n_hits_per_frame=0.0):
    """SURVEY.md 8d per frame: pyramid = read the source once + write every level once;
    cascade = read every level once + write the hits."""
    chn = 4 * plan.C * sum(lv.u * lv.v for lv in plan.levels)
    pyr = plan.info.H * plan.info.W * dtype_bytes + chn
    cas = chn + 36 * n_hits_per_frame
    return pyr, cas


def run_gpu(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- CPU baseline first (rank 0 at N=1 only), before this process's CUDA work starts: bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_measure(args, 1, 1)

    from waldboost_b200.build import build
    if local == 0:
        build()            # in-tree nvcc build of libwbg.so if the sources changed (normally a no-op)
    barrier()
    import waldboost_b200 as wb
    from waldboost_b200 import synthetic as S
    from waldboost_b200.engine import get_engine

    B = args.batch
    model = wb.Model.load(MODEL_B)
    if args.profile == "dense":
        model.theta = [-np.inf] * len(model)
    eng = get_engine()

    # ---- synthetic frames, page-locked on the host (seeds 1000.., a set of distinct frames cycled to
    # fill the batch).  Every rank gets the SAME set, rotated by its rank, so the per-GPU work is identical (weak
    # scaling): the cascade's cost depends on the frame content.
    uniq = min(B, args.unique_frames)
    pinned = torch.empty((B, H, W), dtype=torch.uint8, pin_memory=True)
    frames = pinned.numpy()
    base = [S.synthetic_frame(1000 + i, H, W) for i in range(uniq)]
    for i in range(B):
        frames[i] = base[(i + rank) % uniq]

    plan = model._plan(eng, H, W)
    handle = model._device_model()
    dev = eng.upload_images(frames)
    chns = eng.pyramid(dev, plan)
    hit_cap = eng.default_hit_cap(plan, B)

    def step_device():
        eng.pyramid(dev, plan, out=chns)
        return eng.cascade_launch(handle, plan, chns, B, hit_cap)

    def read_counts(meta, nbytes):
        n_hits, stats, _ = eng._read_meta(meta, nbytes, B, plan.n_levels)
        return n_hits, int(stats[:, 0].sum()), int(stats[:, 1].sum())

    # ---- device-resident throughput: `value`
    for _ in range(max(args.warmup, 3)):
        _, meta, nbytes = step_device()
    n_hits, n_loc, n_weak = read_counts(meta, nbytes)
    hits_truncated = n_hits > hit_cap        # only in the dense profile: every window is a hit; the work is the same
    barrier()
    eng.profile_enable(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        ev0.record()
        for _ in range(args.steps):
            step_device()
        ev1.record()
        barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    prof = eng.profile_read()
    eng.profile_enable(False)
    frames_total = B * args.steps * world
    value = frames_total / (ms_total * 1e-3)

    # ---- end to end through the public API: pinned host frames in, boxes out.  With several GPUs the per-rank hit
    # lists are gathered on the host of rank 0 inside the timed region (north_star: "only a host-side gather of boxes"):
    # a gloo group next to the NCCL one carries the pickled records, no device collective touches data.
    from waldboost_b200 import sharding
    host_group = dist.new_group(backend="gloo") if world > 1 else None

    gatherer = sharding.HitGatherer(group=host_group, dst=0, presorted=True) if world > 1 else None

    def e2e_step(src):
        """detect on this rank's frames; with several ranks the gather of the step's hit records is handed to the
        gatherer thread (it overlaps the next step's GPU work) and a future is returned in place of the result"""
        out, h = model.detect_batch(src, return_hits=True)
        if world == 1:
            return h, h
        h["frame"] += rank * B                              # global frame indices (detect_batch returns a fresh array)
        return h, gatherer.submit(h, (model.n_loc, model.n_weak))

    def e2e_loop(src, n):
        pending = None
        for _ in range(n):
            h, res = e2e_step(src)
            if world > 1:
                if pending is not None:
                    pending.result()                        # at most one gather in flight behind the GPU
                pending = res
        torch.cuda.synchronize()
        return h, (pending.result()[0] if world > 1 else res)   # every gather has completed when the clock stops

    e2e_loop(frames, 2)
    barrier()
    t0 = time.perf_counter()
    hits, gathered = e2e_loop(frames, args.steps)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e = {"value": frames_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(frames.nbytes),
           "d2h_bytes_per_step": int(hits.nbytes + 8 + 16 * B + 4 * B * plan.n_levels),
           "host_gather": None if world == 1 else f"hit records of {world} ranks gathered on rank 0 (gloo, host memory; the gather of a step overlaps the next step's GPU work, all gathers complete inside the timed region): "
                                                   f"{0 if gathered is None else int(gathered.size)} hits per step"}
    # the same call on ordinary (pageable) NumPy frames: the library stages them through a page-locked buffer
    pageable = np.array(frames)
    e2e_loop(pageable, 1)
    barrier()
    t0 = time.perf_counter()
    e2e_loop(pageable, max(2, args.steps // 3))
    e2e_pg_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e["pageable"] = {"value": B * max(2, args.steps // 3) * world / e2e_pg_s, "unit": UNIT,
                       "note": "np.ndarray frames that are not page-locked (staged through a pinned buffer by the library)"}
    del pageable
    if gatherer is not None:
        gatherer.close()

    # ---- roofline of the two dominant kernels (algorithmic bytes per launch / CUDA-event time per launch)
    peaks, peak_src = {}, "fallback 6650 GB/s (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
        peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    pyr_b, cas_b = algorithmic_bytes(plan, 1, n_hits / B)
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "dram_traffic.json")) as f:
            traffic = json.load(f)
    except Exception:
        pass

    def roof(kind, bytes_per_frame, name, symbol):
        ms, n = prof[kind]
        if n == 0:
            return None
        per_launch_s = ms * 1e-3 / n
        ach = bytes_per_frame * B / per_launch_s / 1e9
        return {"kernel": symbol, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": (traffic[name]["bytes_per_frame"] * B) if name in traffic else None,
                "traffic_source": ("profiles/dram_traffic.json: " + traffic[name].get("source", "") + " -- an ncu --set full capture of a 4-frame launch scaled per frame, NOT measured in this run") if name in traffic else None,
                "algorithmic_bytes_per_launch": int(bytes_per_frame * B),
                "ms_per_launch": per_launch_s * 1e3, "peak_source": peak_src,
                # context from the committed ncu capture (profiles/r02_ncu_full_final.csv), not measured in this run:
                # both kernels are bound by instruction issue, not by HBM
                "ncu": {k: v for k, v in traffic.get(name, {}).items() if k.startswith("ncu_")} or None,
                "share_of_step": ms / n / (ms_total / args.steps)}

    r_pyr = roof("level_kernel", pyr_b, "level_kernel", "level_hist4_u8_kernel")
    r_cas = roof("cascade_kernel", cas_b, "cascade_kernel", "cascade_pool_kernel<MODE_D2,512>")
    dominant = r_cas if (r_cas and r_pyr and r_cas["ms_per_launch"] > r_pyr["ms_per_launch"]) else (r_pyr or r_cas)

    # kernels of libwbg per step: minmax_init, [minmax unless fused into the first octave step: uint8 frames with even
    # height and a width that is a multiple of 8], octave steps, level kernel | cascade kernel, 2 mask scans, emit_hits
    fused_minmax = W % 8 == 0 and H % 2 == 0 and plan.info.n_octaves > 1
    launches_per_step = 1 + (0 if fused_minmax else 1) + (plan.info.n_octaves - 1) + 1 + 4
    n_weak_all = sum_over_ranks(float(n_weak))
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"batch of {B} synthetic 1920x1080 uint8 frames per GPU, 12x12x4 grad_hist model "
                                   f"(shrink 2, n_per_oct 8, smooth 1), 1024 depth-2 stages, {args.profile} thetas "
                                   "(BASELINE configs[1]; image-sharded over GPUs = configs[4])",
                       "profile": args.profile, "batch_per_gpu": B, "unique_frames_per_gpu": uniq,
                       "levels": plan.n_levels, "windows_per_frame": int(plan.info.n_loc),
                       "eval_cost": n_weak / max(n_loc, 1), "hits_per_frame": n_hits / B,
                       "hits_truncated_in_device_loop": hits_truncated,
                       "window_stage_evals_per_s": n_weak_all * args.steps / (ms_total * 1e-3),
                       "l2": f"inputs larger than L2: {frames.nbytes / 1e6:.0f} MB of frames and {chns.numel() * 4 / 1e9:.2f} GB of channels per step",
                       "parallelism": f"image-sharded x{world}, no collective"},
            "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "roofline": dominant, "roofline_pyramid": r_pyr, "roofline_cascade": r_cas,
            "cpu_baseline": cpu, "clocks": clk.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------- config C
MODEL_C = os.path.join(ROOT, "tests", "golden", "configC_model.pb")


def run_config_c(args):
    """BASELINE configs[2]: ONE 3840x2160 frame, 20x20x10 grad_mag + grad_hist(9) model, spread over the GPUs of the box
    in row bands of pyramid levels (waldboost_b200.sharding.assign_bands), the hit records gathered on the host of rank
    0.  A step = detect() of that frame through the public API on every rank: H2D of the frame, the rank's bands of the
    pyramid and of the cascade, D2H of its hits, host gather.  Rank 0 checks the gathered list against the reference's
    own detect() output (tests/golden/configC_detect.npz)."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback")
    torch.cuda.set_device(local)
    host_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        host_group = dist.new_group(backend="gloo")
    import waldboost_b200 as wb
    from waldboost_b200 import sharding, synthetic as S
    from waldboost_b200.engine import cascade_tile, get_engine, plan_geometry

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    model = wb.Model.load(MODEL_C)
    Hc, Wc = 2160, 3840
    pinned = torch.empty((1, Hc, Wc), dtype=torch.uint8, pin_memory=True)
    frame = pinned.numpy()
    frame[0] = S.synthetic_frame(1000, Hc, Wc)
    m, n, C_ = (int(x) for x in model.shape)
    plan = plan_geometry(Hc, Wc, model.channel_opts, model._spec(), m, n)
    TR, _ = cascade_tile(m, n, C_)
    rows = [(lv.win_rows + TR - 1) // TR if lv.win_rows > 0 and lv.win_cols > 0 else 0 for lv in plan.levels]
    cost = [TR * lv.v for lv in plan.levels]
    mine = sharding.assign_bands(rows, cost, world)[rank]

    t_detect = [0.0]

    def step():
        model.reset()
        ta = time.perf_counter()
        _, h = model.detect_batch(frame, return_hits=True, bands=mine)
        t_detect[0] += time.perf_counter() - ta
        return sharding.gather_hits(h, (model.n_loc, model.n_weak), group=host_group, dst=0)

    for _ in range(max(args.warmup, 3)):
        hits, stats = step()
    t_detect[0] = 0.0
    barrier()
    eng = get_engine()
    eng.profile_enable(True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        hits, stats = step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    prof = eng.profile_read()
    eng.profile_enable(False)
    dev_ms = sum(ms for ms, _ in prof.values()) / args.steps        # this rank's two dominant kernels per step
    if world > 1:
        t = torch.tensor([dt, dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, dev_ms = float(t[0].item()), float(t[1].item())
    barrier()
    if rank == 0:
        g = np.load(os.path.join(ROOT, "tests", "golden", "configC_detect.npz"))
        boxes = np.stack([hits["x1"], hits["y1"], hits["x2"], hits["y2"]], axis=1)
        same = bool(np.array_equal(boxes, g["boxes"]) and np.array_equal(hits["score"], g["scores"])
                    and stats == (int(g["n_loc"]), int(g["n_weak"])))
        line = {
            "metric": "3840x2160 frames/sec for Model.detect, one frame spread over the GPUs", "value": args.steps / dt,
            "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "single 3840x2160 uint8 frame, 20x20x10 grad_mag + grad_hist(9) model (shrink 2, n_per_oct 8, smooth 1), "
                                   "256 depth-2 stages, wald thetas (BASELINE configs[2]); pyramid levels split in row bands across the GPUs",
                       "windows": int(plan.info.n_loc), "levels": plan.n_levels, "bands_rank0": len(mine),
                       "parallelism": f"(level, row band) x{world}, host gather of hits, no collective on data",
                       "hits": int(hits.size), "matches_reference_golden": same,
                       "device_ms_per_step_max_over_ranks": dev_ms,      # level kernel + cascade kernel of the slowest rank (CUDA events)
                       "rank0_detect_ms_per_step": 1e3 * t_detect[0] / args.steps,
                       "rank0_gather_and_wait_ms_per_step": 1e3 * (dt - t_detect[0]) / args.steps,
                       "timed_region": "public API per step on every rank: H2D of the 8.3 MB frame, bands of pyramid + cascade, D2H of hits, gloo gather on rank 0"},
            "e2e": {"value": args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(frame.nbytes) * world,
                    "d2h_bytes_per_step": int(hits.nbytes)},
            "gpu_launches": None,
        }
        print(json.dumps(line), flush=True)
        if not same:
            raise SystemExit("config C: the gathered hit list differs from the reference golden")
    if world > 1:
        dist.destroy_process_group()


def run_config_small(args):
    """BASELINE configs[0] (A: one 640x480 frame through Model.detect, the reference's own CPU-runnable case) and
    configs[3] (D: dense scoring of 1000 640x480 images with a 2048-stage depth-4 cascade, the Pool.update pattern), one
    GPU, through the public API with host buffers.  Both use the committed reference-written models and compare a
    result with the reference's own output before timing (tests/golden/configA_detect.npz, configD_scan.npz)."""
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback")
    if int(os.environ.get("RANK", "0")) != 0:
        return                                              # single-GPU configurations: other ranks have nothing to do
    import waldboost_b200 as wb
    from waldboost_b200 import synthetic as S
    golden = os.path.join(ROOT, "tests", "golden")
    if args.config == "A":
        M = wb.Model.load(os.path.join(golden, "configA_model.pb"))
        g = np.load(os.path.join(golden, "configA_detect.npz"))
        frame = S.synthetic_frame(1000, 480, 640)
        M.reset()
        dt = M.detect(frame)
        same = bool(np.array_equal(dt.get(), g["boxes"]) and np.array_equal(dt.get_field("scores"), g["scores"])
                    and (M.n_loc, M.n_weak) == (int(g["n_loc"]), int(g["n_weak"])))
        run, units, what = (lambda: M.detect(frame)), 1, "640x480 frames/sec for Model.detect, one frame at a time (latency)"
        workload = "one 640x480 uint8 frame per step, 12x12x4 grad_hist model, 256 depth-2 stages, wald thetas (BASELINE configs[0])"
        h2d, extra = frame.nbytes, {"hits": len(dt), "eval_cost": M.eval_cost}
    else:
        M = wb.Model.load(os.path.join(golden, "configD_model.pb"))
        g = np.load(os.path.join(golden, "configD_scan.npz"))
        small = S.synthetic_frame(1003, 200, 260)
        M.reset()
        got = [h for _, _, (_, _, h) in M.scan_channels(small)]
        same = bool(len(got) == int(g["n_levels"]) and all(np.array_equal(h, g[f"{k}/h"]) for k, h in enumerate(got))
                    and (M.n_loc, M.n_weak) == (int(g["n_loc"]), int(g["n_weak"])))
        base = [S.synthetic_frame(1000 + i, 480, 640) for i in range(16)]
        pinned = torch.empty((1000, 480, 640), dtype=torch.uint8, pin_memory=True)
        frames = pinned.numpy()
        for i in range(1000):
            frames[i] = base[i % 16]
        M.reset()
        run, units, what = (lambda: M.detect_batch(frames, return_hits=True)), 1000, "640x480 images/sec, dense scoring with a 2048-stage depth-4 cascade"
        workload = ("1000 synthetic 640x480 uint8 images per step (16 distinct ones cycled), 12x12x4 grad_hist model, 2048 depth-4 stages, "
                    "wald thetas (BASELINE configs[3]); every image's surviving windows and scores come back to the host")
        h2d, extra = frames.nbytes, {}
    for _ in range(max(args.warmup, 3)):
        out = run()
    torch.cuda.synchronize()
    steps = args.steps if args.config == "D" else max(args.steps, 50)
    with ClockSampler(0) as clk:
        t0 = time.perf_counter()
        for _ in range(steps):
            out = run()
        torch.cuda.synchronize()
        dt_s = time.perf_counter() - t0
    if args.config == "D":
        extra = {"hits_per_image": out[1].size / 1000.0, "eval_cost": M.eval_cost}
    line = {"metric": what, "value": units * steps / dt_s, "unit": "frames/s", "n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 3),
            "ms_per_step": 1e3 * dt_s / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": dict({"workload": workload, "matches_reference_golden": same,
                            "timed_region": "public API per step: host frames in (H2D inside), boxes / hit records out (D2H inside)"}, **extra),
            "e2e": {"value": units * steps / dt_s, "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": None},
            "gpu_launches": None, "clocks": clk.summary()}
    print(json.dumps(line), flush=True)
    if not same:
        raise SystemExit(f"config {args.config}: the result differs from the reference golden")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="frames per GPU per step")
    ap.add_argument("--unique-frames", type=int, default=64, help="distinct synthetic frames per GPU (cycled to fill the batch)")
    ap.add_argument("--profile", choices=["wald", "dense"], default="wald")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--cpu-workers", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-single-thread", action="store_true", help="skip the single-core sample of the CPU arm")
    ap.add_argument("--cpu-kind", choices=["auto", "port", "reference"], default="auto",
                    help="CPU arm: the unmodified reference (needs /root/reference), its NumPy restatement, or whichever is available")
    ap.add_argument("--config", choices=["A", "B", "C", "D"], default="B", help="B: batch of 1080p frames (default, the headline); "
                    "C: one 3840x2160 frame spread over the GPUs in row bands; A: one 640x480 frame (latency); "
                    "D: dense scoring of 1000 640x480 images with a 2048-stage depth-4 cascade")
    args = ap.parse_args()
    if args.config == "C" and args.impl != "reference":
        return run_config_c(args)
    if args.config in ("A", "D") and args.impl != "reference":
        return run_config_small(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    main()
