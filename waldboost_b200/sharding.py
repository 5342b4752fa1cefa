"""Multi-GPU partitioning of the detect() path: frames are independent, so a batch is sharded by image
across ranks (one process per GPU) and the per-rank hit lists are gathered on the host; a single huge frame is
sharded by pyramid level (every level depends only on the original image, reference channels.py:95-101,125-132).
There is no data-path collective: the only communication is the host-side gather of the hit records (and the sum of
the n_loc / n_weak counters), done through `torch.distributed`'s object gather on whatever backend the group has.
"""
import numpy as np


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of n_items for `rank` of `world`; sizes differ by at most one."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} of world {world}")
    base, extra = divmod(int(n_items), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def assign_levels(costs, world):
    """Greedy longest-processing-time assignment of pyramid levels to ranks.  `costs[l]` ~ work of level l
    (channel pixels u*v); returns a list of sorted level-index lists, one per rank."""
    order = sorted(range(len(costs)), key=lambda l: (-costs[l], l))
    load = [0.0] * world
    out = [[] for _ in range(world)]
    for l in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(l)
        load[r] += costs[l]
    return [sorted(x) for x in out]


def assign_bands(level_tile_rows, level_row_cost, world):
    """Split one frame's pyramid into `world` contiguous pieces of (nearly) equal cost.

    The work is laid out as one sequence -- the window-tile rows of level 0, then those of level 1, ... -- where a tile
    row of level l costs level_row_cost[l] (~ its channel pixels), and the sequence is cut into `world` runs.  A rank
    therefore gets at most ONE contiguous band of any level (what wbg_plan_create_bands takes), and the pieces differ
    by at most one tile row of the largest level.  Returns, per rank, a list of (level, first tile row, tile rows)."""
    items = [(l, r, float(level_row_cost[l])) for l, n in enumerate(level_tile_rows) for r in range(int(n))]
    total = sum(c for *_, c in items)
    out = [[] for _ in range(world)]
    if not items:
        return out
    acc, k = 0.0, 0
    for l, r, c in items:
        # the item goes to the rank whose cost interval contains its midpoint
        k = min(world - 1, int((acc + 0.5 * c) * world / total)) if total > 0 else 0
        if out[k] and out[k][-1][0] == l and out[k][-1][1] + out[k][-1][2] == r:
            out[k][-1] = (l, out[k][-1][1], out[k][-1][2] + 1)
        else:
            out[k].append((l, r, 1))
        acc += c
    return out


def normalise_hits(hits):
    """Order hit records like the reference's output: (frame, level, r, c) ascending (model.py:173-179 per frame)."""
    if hits.size == 0:
        return hits
    order = np.lexsort((hits["c"], hits["r"], hits["level"], hits["frame"]))
    return hits[order]


def gather_hits(local_hits, local_stats, group=None, dst=0, presorted=False):
    """Host-side gather of per-rank hit records (frame indices already global) and (n_loc, n_weak) counters.
    Returns (hits ordered by (frame, level, r, c), (n_loc, n_weak)) on rank `dst`, (None, None) elsewhere.
    `presorted`: every rank's list is already in that order and the ranks hold ascending frame ranges (image
    sharding), so concatenating in rank order needs no sort."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return (local_hits if presorted else normalise_hits(local_hits)), tuple(int(x) for x in local_stats)
    import torch
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    backend = dist.get_backend(group)
    if backend != "gloo":
        # device-side backends would move the records through GPU memory: use the (slower) object gather there
        payload = (local_hits.tobytes(), local_hits.dtype.descr, tuple(int(x) for x in local_stats))
        bucket = [None] * world if rank == dst else None
        dist.gather_object(payload, bucket, dst=dst, group=group)
        if rank != dst:
            return None, None
        parts = [np.frombuffer(b, dtype=np.dtype(d)) for b, d, _ in bucket]
        stats = (sum(s[0] for *_, s in bucket), sum(s[1] for *_, s in bucket))
    else:
        item = local_hits.dtype.itemsize
        hdr_np = np.array([int(local_hits.size), int(local_stats[0]), int(local_stats[1])], np.int64)
        raw = np.ascontiguousarray(local_hits).view(np.uint8).reshape(-1)
        # host memory, two collectives, no pickling: the counts (and counters) of every rank, then the records padded to
        # the longest list.  (Short lists in ONE fixed-size all_gather was tried for config C: the same latency on 2 and
        # 4 ranks, 0.4 ms slower on 8 -- gloo's ring all_gather pays per step for the payload.)
        hdr = torch.from_numpy(hdr_np)
        all_t = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(all_t, hdr, group=group)
        all_hdr = [t.numpy() for t in all_t]
        counts = [int(h[0]) for h in all_hdr]
        cap = max(max(counts), 1) * item
        buf = torch.zeros(cap, dtype=torch.uint8)
        if raw.size:
            buf[:raw.size] = torch.from_numpy(raw)
        bucket = [torch.empty(cap, dtype=torch.uint8) for _ in range(world)] if rank == dst else None
        dist.gather(buf, bucket, dst=dst, group=group)
        if rank != dst:
            return None, None
        parts = [b.numpy()[:c * item].view(local_hits.dtype) for b, c in zip(bucket, counts)]
        stats = (sum(int(h[1]) for h in all_hdr), sum(int(h[2]) for h in all_hdr))
    hits = np.concatenate(parts) if parts else local_hits
    return (hits if presorted else normalise_hits(hits)), stats


class HitGatherer:
    """Pipelined host gather: `submit()` hands a rank's hit records to ONE background thread that runs `gather_hits`
    on them, so the collective of step i overlaps the GPU work of step i+1.  Every rank submits in the same order and
    the single worker thread keeps that order, which is all the process group needs.  `submit` returns a
    concurrent.futures.Future whose result is what gather_hits returns."""

    def __init__(self, group=None, dst=0, presorted=False):
        from concurrent.futures import ThreadPoolExecutor
        self.group, self.dst, self.presorted = group, dst, presorted
        self._pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="wbg-gather")

    def submit(self, local_hits, local_stats):
        return self._pool.submit(gather_hits, local_hits, tuple(int(x) for x in local_stats), self.group, self.dst, self.presorted)

    def close(self):
        self._pool.shutdown(wait=True)


def detect_sharded(detect_fn, frames, group=None, dst=0):
    """Image-sharded detect over the ranks of `group`.

    `frames` is the FULL [B,H,W] batch (every rank holds or can index it); `detect_fn(frames_shard)` runs the local
    detector and returns (hit records with shard-local `frame` indices, (n_loc, n_weak)) -- on a GPU rank this is
    `Model.detect_batch(..., return_hits=True)` on that rank's device.  Returns (hits, stats) on `dst`."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    lo, hi = shard_range(len(frames), rank, world)
    if hi > lo:
        hits, stats = detect_fn(frames[lo:hi])
        hits = hits.copy()
        hits["frame"] += lo
    else:
        from ._native import HIT_DTYPE
        hits, stats = np.empty(0, HIT_DTYPE), (0, 0)
    return gather_hits(hits, stats, group, dst)


def detect_level_sharded(detect_levels_fn, level_costs, group=None, dst=0):
    """One huge frame spread over the ranks of `group` by pyramid level (BASELINE config C).

    `level_costs[l]` ~ work of level l (channel pixels); `detect_levels_fn(level_ids)` runs the local detector on the
    given levels and returns (hit records with global `level` indices, (n_loc, n_weak)) -- on a GPU rank
    `Model.detect_batch(frame[None], return_hits=True, levels=level_ids)`.  Returns (hits, stats) on `dst`."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    mine = assign_levels(level_costs, world)[rank]
    if mine:
        hits, stats = detect_levels_fn(mine)
    else:
        from ._native import HIT_DTYPE
        hits, stats = np.empty(0, HIT_DTYPE), (0, 0)
    return gather_hits(hits, stats, group, dst)
