"""Host-side driver of libwbg: plans, device buffers and launches.  PyTorch is used only as the device-memory and
stream carrier; every computation is a C-ABI call into the CUDA library (include/wbg.h).  No CPU fallback."""
import ctypes as C
import os
import threading

import numpy as np

from . import _native as N

_engines = {}
_lock = threading.Lock()


def _torch():
    import torch
    return torch


def get_engine(device=None):
    """Engine bound to one CUDA device (default: the current one).  Raises without a GPU."""
    torch = _torch()
    if not torch.cuda.is_available():
        raise RuntimeError("waldboost_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    with _lock:
        eng = _engines.get(idx)
        if eng is None:
            eng = _engines[idx] = Engine(idx)
    return eng


def make_channel_opts(channel_opts, spec, max_levels=0):
    """dict + resolved channel spec -> wbg_channel_opts.  cos/sin exactly as reference channels.py:43-46."""
    o = N.ChannelOpts()
    o.shrink = int(channel_opts["shrink"])
    o.n_per_oct = int(channel_opts["n_per_oct"])
    o.smooth = int(channel_opts["smooth"])
    o.kind = spec["kind"]
    o.n_bins = int(spec["n_bins"])
    o.full = 1 if spec["full"] else 0
    o.bias = float(spec["bias"])
    o.norm = int(spec["norm"] or 0)
    o.eps = float(spec["eps"])
    o.max_levels = int(max_levels)
    if o.n_bins > N.WBG_MAX_BINS:
        raise ValueError(f"n_bins must be <= {N.WBG_MAX_BINS}")
    if o.n_bins > 0 and spec["kind"] in (N.WBG_CH_GRAD_HIST, N.WBG_CH_GRAD_MAG_HIST):
        max_theta = 2 * np.pi if spec["full"] else np.pi
        theta = np.linspace(0, max_theta, o.n_bins + 1)
        cs, sn = np.cos(theta[:-1]), np.sin(theta[:-1])
        for i in range(o.n_bins):
            o.cos_t[i] = cs[i]
            o.sin_t[i] = sn[i]
    return o


def _opts_key(channel_opts, spec, max_levels):
    return (int(channel_opts["shrink"]), int(channel_opts["n_per_oct"]), int(channel_opts["smooth"]), spec["kind"],
            int(spec["n_bins"]), bool(spec["full"]), float(spec["bias"]), int(spec["norm"] or 0), float(spec["eps"]),
            int(max_levels))


class Plan:
    """Pyramid geometry for one (H, W, channel options, window) -- wraps a wbg_plan handle."""

    def __init__(self, H, W, copts, win_m, win_n, device_tables=True, level_ids=None, bands=None):
        L = N.lib()
        self.handle = C.c_void_p()
        if bands is not None:
            flat = [int(x) for b in sorted(bands) for x in b]           # triples (level, first tile row, tile rows)
            arr = (C.c_int32 * max(len(flat), 1))(*flat)
            code = L.wbg_plan_create_bands(H, W, C.byref(copts), win_m, win_n, 1 if device_tables else 0, arr, len(flat) // 3,
                                           C.byref(self.handle))
        elif level_ids is None:
            code = L.wbg_plan_create(H, W, C.byref(copts), win_m, win_n, 1 if device_tables else 0, C.byref(self.handle))
        else:
            ids = sorted(set(int(x) for x in level_ids))
            arr = (C.c_int32 * max(len(ids), 1))(*ids)
            code = L.wbg_plan_create_levels(H, W, C.byref(copts), win_m, win_n, 1 if device_tables else 0, arr, len(ids),
                                            C.byref(self.handle))
        if code == N.WBG_EINVAL and "Shrink factor" in N.last_error():
            raise AssertionError(N.last_error())          # reference channels.py:120 is an assert
        N.check(code)
        self.info = N.PlanInfo()
        N.check(L.wbg_plan_get_info(self.handle, C.byref(self.info)))
        n = self.info.n_levels
        arr = (N.Level * max(n, 1))()
        N.check(L.wbg_plan_get_levels(self.handle, arr, max(n, 1)))
        self.levels = [arr[i] for i in range(n)]
        self.n_levels = n
        self.C = self.info.channels
        self.chn_floats = self.info.chn_floats
        self.scales = [lv.scale for lv in self.levels]

    def __del__(self):
        try:
            if self.handle:
                N.lib().wbg_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def plan_geometry(H, W, channel_opts, spec, win_m=0, win_n=0, max_levels=0, level_ids=None, bands=None):
    """Host-only plan (no GPU needed): level sizes, offsets and window counts."""
    return Plan(H, W, make_channel_opts(channel_opts, spec, max_levels), win_m, win_n, device_tables=False, level_ids=level_ids, bands=bands)


def cascade_tile(win_m, win_n, channels):
    """(rows, columns) of the cascade kernel's window tile for a win_m x win_n x channels model (host arithmetic)."""
    tr, tc = C.c_int32(), C.c_int32()
    N.check(N.lib().wbg_cascade_tile(int(win_m), int(win_n), int(channels), C.byref(tr), C.byref(tc)))
    return tr.value, tc.value


class ModelHandle:
    def __init__(self, shape, trees, thetas):
        """trees: list of objects with feature/threshold/left/right/prediction arrays (training.DTree)."""
        m, n, ch = (int(x) for x in shape)
        T = len(trees)
        Nn = max([len(t.left) for t in trees] + [1])
        n_nodes = np.zeros(max(T, 1), np.int32)
        feature = np.zeros((max(T, 1), Nn, 3), np.uint8)
        threshold = np.zeros((max(T, 1), Nn), np.float32)
        left = np.full((max(T, 1), Nn), -1, np.int8)
        right = np.full((max(T, 1), Nn), -1, np.int8)
        prediction = np.zeros((max(T, 1), Nn), np.float32)
        for t, tr in enumerate(trees):
            k = len(tr.left)
            n_nodes[t] = k
            feature[t, :k] = np.asarray(tr.feature, np.uint8).reshape(-1, 3)
            threshold[t, :k] = tr.threshold
            left[t, :k] = tr.left
            right[t, :k] = tr.right
            prediction[t, :k] = tr.prediction
        # NumPy compares the float32 scores with a (weak) Python-float theta in float32 (model.py:255)
        with np.errstate(over="ignore"):
            theta = np.asarray([float(x) for x in thetas] + ([] if T else [0.0]), np.float64).astype(np.float32)
        d = N.ModelDesc()
        d.win_m, d.win_n, d.channels, d.n_stages, d.max_nodes = m, n, ch, T, Nn
        keep = (n_nodes, feature, threshold, left, right, prediction, theta)
        d.n_nodes, d.feature, d.threshold = n_nodes.ctypes.data, feature.ctypes.data, threshold.ctypes.data
        d.left, d.right, d.prediction, d.theta = left.ctypes.data, right.ctypes.data, prediction.ctypes.data, theta.ctypes.data
        self.handle = C.c_void_p()
        code = N.lib().wbg_model_create(C.byref(d), C.byref(self.handle))
        del keep
        if code == N.WBG_EINVAL:
            raise ValueError(N.last_error())
        N.check(code)
        self.shape = (m, n, ch)
        self.T = T

    def __del__(self):
        try:
            if self.handle:
                N.lib().wbg_model_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def _check_windows(rs, cs, u, v, m, n):
    """Window origins for the explicit-window entry points.  The reference gathers X[rs + r, cs + c] with NumPy fancy
    indexing (training.py:92, samples.py:14-43): an origin whose window leaves the map raises IndexError there, and the
    device kernels do not bounds-check, so the same error is raised here before anything is launched."""
    rs = np.ascontiguousarray(rs, np.int64).ravel()
    cs = np.ascontiguousarray(cs, np.int64).ravel()
    if rs.size != cs.size:
        raise ValueError("Sizes of 'rs' and 'cs' must match")
    if rs.size and (rs.min() < 0 or cs.min() < 0 or rs.max() + m > u or cs.max() + n > v):
        raise IndexError(f"window origin outside the {u} x {v} channel map for a {m} x {n} window")
    return rs.astype(np.int32), cs.astype(np.int32)


class Engine:
    def __init__(self, device_index):
        torch = _torch()
        self.torch = torch
        self.device = torch.device("cuda", device_index)
        self.lib = N.lib()
        self._plans = {}
        self._bufs = {}
        self._pinned = {}

    # ------------------------------------------------------------------------------------------- resources
    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def buffer(self, name, nbytes):
        """Grow-only cached device byte buffer (256-byte aligned by the torch allocator)."""
        t = self._bufs.get(name)
        if t is None or t.numel() < nbytes:
            t = None
            self._bufs.pop(name, None)
            t = self.torch.empty(max(int(nbytes), 256), dtype=self.torch.uint8, device=self.device)
            self._bufs[name] = t
        return t

    def pinned(self, name, nbytes):
        t = self._pinned.get(name)
        if t is None or t.numel() < nbytes:
            t = self.torch.empty(max(int(nbytes), 256), dtype=self.torch.uint8, pin_memory=True)
            self._pinned[name] = t
        return t

    def plan(self, H, W, channel_opts, spec, win_m=0, win_n=0, max_levels=0, level_ids=None, bands=None):
        lv_key = None if level_ids is None else tuple(sorted(set(int(x) for x in level_ids)))
        bd_key = None if bands is None else tuple(sorted(tuple(int(x) for x in b) for b in bands))
        key = (H, W, win_m, win_n, lv_key, bd_key) + _opts_key(channel_opts, spec, max_levels)
        p = self._plans.get(key)
        if p is None:
            with self.torch.cuda.device(self.device):
                p = Plan(H, W, make_channel_opts(channel_opts, spec, max_levels), win_m, win_n, level_ids=lv_key, bands=bd_key)
            if len(self._plans) > 64:
                self._plans.clear()
            self._plans[key] = p
        return p

    # ------------------------------------------------------------------------------------------- pyramid
    @staticmethod
    def image_dtype(dtype):
        if dtype == np.uint8:
            return N.WBG_U8
        if dtype == np.float32:
            return N.WBG_F32
        raise TypeError(f"image dtype {dtype} is not supported on the GPU path (uint8 and float32 only; no CPU fallback)")

    def upload_images(self, images, slot=""):
        """numpy [B,H,W] (uint8 / float32) -> device tensor.  Page-locked input (e.g. a NumPy view of a pinned torch
        tensor) is copied straight to the device; pageable input goes through a cached pinned staging buffer."""
        torch = self.torch
        images = np.ascontiguousarray(images)
        self.image_dtype(images.dtype)
        nbytes = images.nbytes
        tdt = torch.uint8 if images.dtype == np.uint8 else torch.float32
        dev = self.buffer("img" + slot, nbytes)[:nbytes].view(tdt).view(images.shape)
        src = torch.from_numpy(images) if images.flags.writeable else None
        if src is not None and src.is_pinned():
            dev.copy_(src, non_blocking=True)
            return dev
        stage = self.pinned("img" + slot, nbytes)
        stage_np = stage.numpy()[:nbytes].view(images.dtype).reshape(images.shape)
        np.copyto(stage_np, images)
        dev.copy_(stage[:nbytes].view(tdt).view(images.shape), non_blocking=True)
        return dev

    def pyramid(self, img_dev, plan, out=None, slot=""):
        """img_dev: device tensor [B,H,W] uint8/float32 -> chns device tensor [B, chn_floats] float32."""
        torch = self.torch
        B = int(img_dev.shape[0])
        dt = N.WBG_U8 if img_dev.dtype == torch.uint8 else N.WBG_F32
        assert img_dev.is_contiguous() and img_dev.shape[1] == plan.info.H and img_dev.shape[2] == plan.info.W
        if out is None:
            out = self.buffer("chns" + slot, 4 * B * max(plan.chn_floats, 1))[:4 * B * plan.chn_floats].view(torch.float32).view(B, plan.chn_floats)
        wsb = self.lib.wbg_pyramid_workspace_bytes(plan.handle, dt, B)
        ws = self.buffer("pyr_ws" + slot, wsb)
        with torch.cuda.device(self.device):
            N.check(self.lib.wbg_channel_pyramid(plan.handle, C.c_void_p(img_dev.data_ptr()), dt, B,
                                                 C.c_void_p(out.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(),
                                                 self._stream()))
        return out

    def split_levels(self, chns_row_np, plan):
        """one frame's channel block (numpy, host) -> list of (u,v,C) views."""
        out = []
        for lv in plan.levels:
            n = lv.u * lv.v * plan.C
            out.append(chns_row_np[lv.chn_off:lv.chn_off + n].reshape(lv.u, lv.v, plan.C))
        return out

    def channel_levels(self, image, channel_opts, spec, max_levels=0):
        """[(chns, scale)] for one image -- host copies, like the reference generator's yields."""
        image = np.ascontiguousarray(image)
        H, W = image.shape
        plan = self.plan(H, W, channel_opts, spec, 0, 0, max_levels)
        if plan.n_levels == 0:
            return []
        dev = self.upload_images(image[None])
        chns = self.pyramid(dev, plan)
        host = chns[0].cpu().numpy()
        # the reference's integer channel functions yield uint8 maps (fpga/channels.py:52,63)
        cast = (lambda a: a.astype(np.uint8)) if spec.get("integer") else (lambda a: a.copy())
        return [(cast(lv), s) for lv, s in zip(self.split_levels(host, plan), plan.scales)]

    # ------------------------------------------------------------------------------------------- cascade
    def _meta(self, B, n_levels, slot=""):
        """one device buffer holding [n_hits i64][stats u64 x 2B][level_counts i32 x B*L] -> sub-pointers."""
        nbytes = 8 + 16 * B + 4 * B * n_levels
        t = self.buffer("meta" + slot, nbytes)
        base = t.data_ptr()
        return t, nbytes, base, base + 8, base + 8 + 16 * B

    def _read_meta(self, t, nbytes, B, n_levels):
        host = self.pinned("meta", nbytes)
        host[:nbytes].copy_(t[:nbytes], non_blocking=True)
        self.torch.cuda.current_stream(self.device).synchronize()
        raw = host.numpy()[:nbytes]
        n_hits = int(raw[:8].view(np.int64)[0])
        stats = raw[8:8 + 16 * B].view(np.uint64).reshape(B, 2).copy()
        counts = raw[8 + 16 * B:nbytes].view(np.int32).reshape(B, n_levels).copy()
        return n_hits, stats, counts

    def _read_hits(self, hits_t, n):
        if n == 0:
            return np.empty(0, N.HIT_DTYPE)
        nb = n * N.HIT_DTYPE.itemsize
        host = self.pinned("hits", nb)
        host[:nb].copy_(hits_t[:nb], non_blocking=True)
        self.torch.cuda.current_stream(self.device).synchronize()
        return host.numpy()[:nb].view(N.HIT_DTYPE).copy()

    def cascade_launch(self, model_handle, plan, chns, B, hit_cap, slot=""):
        """Enqueue the cascade over every level of B frames on the current stream (no synchronisation).
        Returns (hits device buffer, meta device buffer, meta bytes)."""
        lib = self.lib
        ws = self.buffer("cas_ws" + slot, lib.wbg_cascade_workspace_bytes(plan.handle, B))
        hits_t = self.buffer("hits" + slot, hit_cap * N.HIT_DTYPE.itemsize)
        meta, nbytes, p_nhits, p_stats, p_counts = self._meta(B, plan.n_levels, slot)
        with self.torch.cuda.device(self.device):
            code = lib.wbg_cascade_scan(model_handle.handle, plan.handle, C.c_void_p(chns.data_ptr()), B,
                                        C.c_void_p(hits_t.data_ptr()), hit_cap, C.c_void_p(p_counts),
                                        C.c_void_p(p_stats), C.c_void_p(p_nhits), C.c_void_p(ws.data_ptr()),
                                        ws.numel(), self._stream())
        if code == N.WBG_EINVAL and "Invalid number of channels" in N.last_error():
            raise AssertionError(N.last_error())       # reference model.py:238 is an assert
        N.check(code)
        return hits_t, meta, nbytes

    def default_hit_cap(self, plan, B):
        return int(min(max(plan.info.n_loc * B, 1), 1 << 20))

    def cascade(self, model_handle, plan, chns, B, hit_cap=None):
        """Run the cascade over every level of B frames.  Returns (hits, level_counts [B,L], stats [B,2])."""
        if hit_cap is None:
            hit_cap = self.default_hit_cap(plan, B)
        while True:
            hits_t, meta, nbytes = self.cascade_launch(model_handle, plan, chns, B, hit_cap)
            # the counters and an optimistic prefix of the hit list come back with ONE synchronisation (the list is short
            # in practice); a longer list costs a second copy
            pre = min(hit_cap, 2048) * N.HIT_DTYPE.itemsize
            host = self.pinned("meta+hits", nbytes + pre)
            host[:nbytes].copy_(meta[:nbytes], non_blocking=True)
            host[nbytes:nbytes + pre].copy_(hits_t[:pre], non_blocking=True)
            self.torch.cuda.current_stream(self.device).synchronize()
            raw = host.numpy()
            n_hits = int(raw[:8].view(np.int64)[0])
            stats = raw[8:8 + 16 * B].view(np.uint64).reshape(B, 2).copy()
            counts = raw[8 + 16 * B:nbytes].view(np.int32).reshape(B, plan.n_levels).copy()
            if n_hits <= hit_cap:
                nb = n_hits * N.HIT_DTYPE.itemsize
                if nb <= pre:
                    return raw[nbytes:nbytes + nb].view(N.HIT_DTYPE).copy(), counts, stats
                return self._read_hits(hits_t, n_hits), counts, stats
            hit_cap = n_hits  # WBG_ECAP semantics: only the first hit_cap were stored -> re-run with room for all

    def run_frames(self, model_handle, plan, images, chunk=None):
        """detect() over host frames [B,H,W] as a two-slot pipeline: the frames go to the device in chunks, and the
        host->device copy of chunk k+1 (copy engine) overlaps the pyramid + cascade kernels of chunk k (two streams,
        one buffer set each).  Returns (hits with batch-global frame indices, level_counts [B,L], stats [B,2])."""
        torch = self.torch
        B = int(images.shape[0])
        if chunk is None:
            # about eight 1080p frames' worth of pixels per chunk: long enough grids for small frames (a chunk of eight
            # 640x480 frames is only ~7 waves of cascade tiles), short enough to overlap the copies of large ones
            auto = max(1, min(256, round(8 * 1920 * 1080 / max(1, int(images.shape[1]) * int(images.shape[2])))))
            chunk = int(os.environ.get("WBG_PIPE_CHUNK", auto))
        if B <= chunk:
            dev = self.upload_images(images)
            chns = self.pyramid(dev, plan)
            return self.cascade(model_handle, plan, chns, B)
        if not hasattr(self, "_pipe_streams"):
            self._pipe_streams = [torch.cuda.Stream(device=self.device) for _ in range(2)]
        cur = torch.cuda.current_stream(self.device)
        # a short first chunk gets the kernels going while the rest of the batch is still being copied; later chunks
        # grow (their copies hide behind ever longer kernels), so most of the batch runs in a few long launches
        grow = int(os.environ.get("WBG_PIPE_GROW", "4"))
        spans, lo, size = [], 0, max(1, chunk // 4)
        while lo < B:
            spans.append((lo, min(lo + size, B)))
            lo += size
            size = chunk if size < chunk else min(2 * size, grow * chunk)
        jobs = []
        for i, (lo, hi) in enumerate(spans):
            slot = str(i & 1)
            st = self._pipe_streams[i & 1]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                n = hi - lo
                dev = self.upload_images(images[lo:hi], slot)
                chns = self.pyramid(dev, plan, slot=slot)
                cap = self.default_hit_cap(plan, n)
                hits_t, meta, nbytes = self.cascade_launch(model_handle, plan, chns, n, cap, slot)
                host_meta = self.pinned("meta" + slot, nbytes)        # two slots: chunk i-1 is read before chunk i+1 is issued
                host_meta[:nbytes].copy_(meta[:nbytes], non_blocking=True)
                # the hit list is small in practice: copy an optimistic prefix right away, the rest on demand
                pre = min(cap, 4096) * N.HIT_DTYPE.itemsize
                host_hits = self.pinned("hits" + slot, pre)
                host_hits[:pre].copy_(hits_t[:pre], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(st)
            jobs.append([lo, hi, slot, st, ev, host_meta, nbytes, host_hits, pre, cap])
            if i >= 1:
                # slot buffers are reused two chunks later: finish reading chunk i-1 before chunk i+1 overwrites them
                self._collect(jobs[i - 1], model_handle, plan, images)
        self._collect(jobs[-1], model_handle, plan, images)
        cur.wait_stream(self._pipe_streams[0])
        cur.wait_stream(self._pipe_streams[1])
        hits = np.concatenate([j[-1][0] for j in jobs]) if jobs else np.empty(0, N.HIT_DTYPE)
        counts = np.concatenate([j[-1][1] for j in jobs], axis=0)
        stats = np.concatenate([j[-1][2] for j in jobs], axis=0)
        return hits, counts, stats

    def _collect(self, job, model_handle, plan, images):
        """wait for one pipeline chunk and read its results back (appends the result tuple to the job record)."""
        lo, hi, slot, st, ev, host_meta, nbytes, host_hits, pre, cap = job[:10]
        if len(job) > 10:
            return
        n = hi - lo
        ev.synchronize()
        raw = host_meta.numpy()[:nbytes]
        n_hits = int(raw[:8].view(np.int64)[0])
        stats = raw[8:8 + 16 * n].view(np.uint64).reshape(n, 2).copy()
        counts = raw[8 + 16 * n:nbytes].view(np.int32).reshape(n, plan.n_levels).copy()
        nb = n_hits * N.HIT_DTYPE.itemsize
        if n_hits > cap:
            # more survivors than the buffer holds: redo this chunk on its stream with room for all of them
            with self.torch.cuda.stream(st):
                dev = self.upload_images(images[lo:hi], slot)
                chns = self.pyramid(dev, plan, slot=slot)
                hits, counts, stats = self.cascade(model_handle, plan, chns, n, hit_cap=n_hits)
        elif nb <= pre:
            hits = host_hits.numpy()[:nb].view(N.HIT_DTYPE).copy()
        else:
            with self.torch.cuda.stream(st):
                hits = self._read_hits(self._bufs["hits" + slot], n_hits)
        hits["frame"] += lo
        job.append((hits, counts, stats))

    def predict_on_map(self, model_handle, X, hit_cap=None):
        """Model.predict_on_image on one channel map X (numpy or device tensor, (u,v,C) float32)."""
        torch, lib = self.torch, self.lib
        m, n, ch = model_handle.shape
        if isinstance(X, np.ndarray):
            X = torch.from_numpy(np.ascontiguousarray(X, np.float32)).to(self.device)
        u, v = int(X.shape[0]), int(X.shape[1])
        n_loc = max(u - m, 0) * max(v - n, 0)
        if hit_cap is None:
            hit_cap = int(min(max(n_loc, 1), 1 << 20))
        wsb = lib.wbg_predict_workspace_bytes(u, v, m, n)
        ws = self.buffer("cas_ws", wsb)
        if X.numel() == 0:
            X = torch.zeros(4, dtype=torch.float32, device=self.device)
        while True:
            hits_t = self.buffer("hits", hit_cap * N.HIT_DTYPE.itemsize)
            meta, nbytes, p_nhits, p_stats, _ = self._meta(1, 1)
            with torch.cuda.device(self.device):
                N.check(lib.wbg_predict_on_image(model_handle.handle, C.c_void_p(X.data_ptr()), u, v,
                                                 C.c_void_p(hits_t.data_ptr()), hit_cap, C.c_void_p(p_stats),
                                                 C.c_void_p(p_nhits), C.c_void_p(ws.data_ptr()), ws.numel(), self._stream()))
            n_hits, stats, _ = self._read_meta(meta, nbytes, 1, 1)
            if n_hits <= hit_cap:
                return self._read_hits(hits_t, n_hits), stats[0]
            hit_cap = n_hits

    def trace(self, model_handle, X, rs, cs):
        """(leaf [K,T] uint8, score [K] float32) for explicit windows, no rejection (wbg_cascade_trace)."""
        torch = self.torch
        X = np.ascontiguousarray(X, np.float32)
        rs, cs = _check_windows(rs, cs, X.shape[0], X.shape[1], model_handle.shape[0], model_handle.shape[1])
        Xd = torch.from_numpy(X).to(self.device)
        rs_d = torch.from_numpy(rs).to(self.device)
        cs_d = torch.from_numpy(cs).to(self.device)
        K, T = int(rs_d.numel()), model_handle.T
        leaf = torch.zeros((K, max(T, 1)), dtype=torch.uint8, device=self.device)
        score = torch.zeros(K, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            N.check(self.lib.wbg_cascade_trace(model_handle.handle, C.c_void_p(Xd.data_ptr()), int(Xd.shape[0]), int(Xd.shape[1]),
                                               C.c_void_p(rs_d.data_ptr()), C.c_void_p(cs_d.data_ptr()), K,
                                               C.c_void_p(leaf.data_ptr()), C.c_void_p(score.data_ptr()), self._stream()))
        return leaf.cpu().numpy()[:, :T], score.cpu().numpy()

    def predict_samples(self, model_handle, X):
        """Model.predict(X) on (K, m, n, C) float32 samples -> (H [K] float32 with -inf for rejected, mask [K] bool)."""
        torch = self.torch
        Xd = torch.from_numpy(np.ascontiguousarray(X, np.float32)).to(self.device)
        K = int(Xd.shape[0])
        H = torch.empty(max(K, 1), dtype=torch.float32, device=self.device)
        mask = torch.empty(max(K, 1), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            N.check(self.lib.wbg_predict_samples(model_handle.handle, C.c_void_p(Xd.data_ptr()), K, C.c_void_p(H.data_ptr()),
                                                 C.c_void_p(mask.data_ptr()), self._stream()))
        return H.cpu().numpy()[:K], mask.cpu().numpy()[:K].astype(bool)

    def gather_samples(self, X, rs, cs, shape):
        torch = self.torch
        m, n = int(shape[0]), int(shape[1])
        X = np.ascontiguousarray(X, np.float32)
        rs, cs = _check_windows(rs, cs, X.shape[0], X.shape[1], m, n)
        Xd = torch.from_numpy(X).to(self.device)
        ch = int(Xd.shape[2])
        rs_d = torch.from_numpy(rs).to(self.device)
        cs_d = torch.from_numpy(cs).to(self.device)
        K = int(rs_d.numel())
        out = torch.empty((K, m, n, ch), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            N.check(self.lib.wbg_gather_samples(C.c_void_p(Xd.data_ptr()), int(Xd.shape[0]), int(Xd.shape[1]), ch,
                                                C.c_void_p(rs_d.data_ptr()), C.c_void_p(cs_d.data_ptr()), K, m, n,
                                                C.c_void_p(out.data_ptr()), self._stream()))
        return out.cpu().numpy()

    # ------------------------------------------------------------------------------------------- measurement
    def profile_enable(self, on):
        N.check(self.lib.wbg_profile_enable(1 if on else 0))

    def profile_read(self):
        """{kind: (kernel ms since the last read, launches)} from the library's CUDA-event brackets."""
        ms = (C.c_double * len(N.PROF_KINDS))()
        n = (C.c_int64 * len(N.PROF_KINDS))()
        N.check(self.lib.wbg_profile_read(ms, n))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(N.PROF_KINDS)}

    def gradients(self, image):
        """(gx, gy) of a 2-D float32 image (reference channels.py:16-21)."""
        torch = self.torch
        h, w = image.shape
        Xd = torch.from_numpy(np.ascontiguousarray(image, np.float32)).to(self.device)
        buf = torch.zeros((3, max(h * w, 1)), dtype=torch.float32, device=self.device)        # gx | gy | scratch
        if h * w:
            with torch.cuda.device(self.device):
                N.check(self.lib.wbg_gradients(C.c_void_p(Xd.data_ptr()), h, w, C.c_void_p(buf[0].data_ptr()),
                                               C.c_void_p(buf[1].data_ptr()), C.c_void_p(buf[2].data_ptr()), self._stream()))
        host = buf[:2, :h * w].cpu().numpy()
        return host[0].reshape(h, w), host[1].reshape(h, w)

    def separable_convolve(self, image, k0, k1=None):
        """k0 along axis 0, then k1 (default k0) along axis 1 (reference channels.py:24-27); symmetric odd kernels."""
        torch = self.torch
        h, w = image.shape
        Xd = torch.from_numpy(np.ascontiguousarray(image, np.float32)).to(self.device)
        buf = torch.zeros((2, max(h * w, 1)), dtype=torch.float32, device=self.device)        # out | scratch
        k0 = np.ascontiguousarray(k0, np.float32)
        k1 = None if k1 is None else np.ascontiguousarray(k1, np.float32)
        with torch.cuda.device(self.device):
            N.check(self.lib.wbg_separable_convolve(C.c_void_p(Xd.data_ptr()), h, w, k0.ctypes.data_as(C.c_void_p), int(k0.size),
                                                    None if k1 is None else k1.ctypes.data_as(C.c_void_p), 0 if k1 is None else int(k1.size),
                                                    C.c_void_p(buf[0].data_ptr()), C.c_void_p(buf[1].data_ptr()), self._stream()))
        return buf[0, :h * w].cpu().numpy().reshape(h, w)

    def map_primitive(self, fn_name, arr, out_shape):
        torch = self.torch
        a3 = arr.reshape(arr.shape[0], arr.shape[1], -1)
        Xd = torch.from_numpy(np.ascontiguousarray(a3, np.float32)).to(self.device)
        out = torch.zeros(int(np.prod(out_shape)) if len(out_shape) else 1, dtype=torch.float32, device=self.device)
        if Xd.numel() and out.numel():
            with torch.cuda.device(self.device):
                N.check(getattr(self.lib, fn_name)(C.c_void_p(Xd.data_ptr()), a3.shape[0], a3.shape[1], a3.shape[2],
                                                   C.c_void_p(out.data_ptr()), self._stream()))
        return out.cpu().numpy()[:int(np.prod(out_shape))].reshape(out_shape)
