"""CPU: the N > 1 path (image sharding + host-side gather) with world_size-2 gloo processes.  The GPU
detector is replaced by the oracle as the per-rank `detect_fn`, so this covers exactly the host logic that runs
between the ranks: shard ranges, global frame indices, gather, order normalisation and counter sums."""
import os
import socket
import sys

import numpy as np
import pytest

from waldboost_b200 import sharding


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 64, 4096):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def test_level_assignment_is_balanced_and_complete():
    """config C: one 3840x2160 frame, 72 levels sharded over 8 ranks by channel-pixel cost."""
    import waldboost_b200 as wb
    from waldboost_b200.engine import plan_geometry
    fn = wb.channels.grad_mag_hist
    plan = plan_geometry(2160, 3840, dict(shrink=2, n_per_oct=8, smooth=1, channels=fn), wb.channels.resolve_channels(fn), 20, 20)
    costs = [lv.u * lv.v for lv in plan.levels]
    parts = sharding.assign_levels(costs, 8)
    assert sorted(l for p in parts for l in p) == list(range(72))
    loads = [sum(costs[l] for l in p) for p in parts]
    assert max(loads) == costs[0] or max(loads) <= 1.34 * sum(costs) / 8     # level 0 alone is 16 % of the pixels
    assert sharding.assign_levels(costs, 1) == [list(range(72))]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_detect_fn():
    import wb_oracle as O
    from waldboost_b200 import synthetic as S
    from waldboost_b200._native import HIT_DTYPE
    opts = dict(shrink=2, n_per_oct=4, smooth=1, channels=O.grad_hist)
    frame0 = S.synthetic_frame(1000, 64, 80)
    lo, hi = S.channel_quantiles(next(iter(O.channel_pyramid(frame0, opts)))[0])
    Cs = O.Cascade((12, 12, 4), opts)
    for t in S.random_trees((12, 12, 4), 6, 2, lo, hi, seed=5):
        Cs.append(O.DTree([tuple(f) for f in t.feature], t.threshold, t.left, t.right, t.prediction), -0.2)

    def detect_fn(frames):
        Cs.reset()
        rec = []
        for b, f in enumerate(frames):
            boxes, scores, levels = Cs.detect(f)
            h = np.zeros(scores.size, HIT_DTYPE)
            h["frame"], h["level"], h["score"] = b, levels, scores
            h["x1"], h["y1"], h["x2"], h["y2"] = boxes.T
            # recover (r, c) from the box corner and the level scale is not needed for ordering tests: use ranks
            h["r"] = np.arange(scores.size) // 1000
            h["c"] = np.arange(scores.size) % 1000
            rec.append(h)
        return np.concatenate(rec), (Cs.n_loc, Cs.n_weak)
    return detect_fn


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "oracle"), os.path.join(root, "tests")):
        sys.path.insert(0, p)
    from waldboost_b200 import synthetic as S
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    frames = S.synthetic_frames(5, 64, 80)
    hits, stats = sharding.detect_sharded(_oracle_detect_fn(), frames)
    if rank == 0:
        np.savez(out_path, hits=hits, stats=np.array(stats))
    else:
        assert hits is None and stats is None
    dist.destroy_process_group()


def test_image_sharded_detect_world2_gloo(tmp_path):
    import torch.multiprocessing as mp
    from waldboost_b200 import synthetic as S
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    ref_hits, ref_stats = sharding.detect_sharded(_oracle_detect_fn(), S.synthetic_frames(5, 64, 80))   # single process
    assert ref_hits.size > 0 and np.array_equal(got["hits"], ref_hits)
    assert tuple(got["stats"]) == tuple(ref_stats)
    assert np.all(np.diff(got["hits"]["frame"]) >= 0) and set(got["hits"]["frame"]) <= set(range(5))


def _level_worker(rank, world, port, out_path):
    import torch.distributed as dist
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "oracle"), os.path.join(root, "tests")):
        sys.path.insert(0, p)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    fn, costs = _oracle_level_fn()
    hits, stats = sharding.detect_level_sharded(fn, costs)
    if rank == 0:
        np.savez(out_path, hits=hits, stats=np.array(stats))
    dist.destroy_process_group()


def _oracle_level_fn():
    """per-rank detector for level sharding: the oracle restricted to a level subset of one frame."""
    import wb_oracle as O
    from waldboost_b200 import synthetic as S
    from waldboost_b200._native import HIT_DTYPE
    opts = dict(shrink=2, n_per_oct=4, smooth=1, channels=O.grad_hist)
    frame = S.synthetic_frame(1001, 120, 160)
    maps = [c for c, _ in O.channel_pyramid(frame, opts)]
    lo, hi = S.channel_quantiles(maps[0])
    Cs = O.Cascade((12, 12, 4), opts)
    for t in S.random_trees((12, 12, 4), 6, 2, lo, hi, seed=5):
        Cs.append(O.DTree([tuple(f) for f in t.feature], t.threshold, t.left, t.right, t.prediction), -0.2)

    def fn(level_ids):
        Cs.reset()
        rec = []
        for lvl, (chns, scale) in zip(sorted(level_ids), Cs.channels(frame, level_ids)):
            r, c, h = Cs.predict_on_image(chns)
            hh = np.zeros(r.size, HIT_DTYPE)
            hh["level"], hh["r"], hh["c"], hh["score"] = lvl, r, c, h
            rec.append(hh)
        return (np.concatenate(rec) if rec else np.empty(0, HIT_DTYPE)), (Cs.n_loc, Cs.n_weak)
    return fn, [m.shape[0] * m.shape[1] for m in maps]


def test_level_sharded_detect_world2_gloo(tmp_path):
    """config C's partitioning: the levels of ONE frame spread over 2 ranks == the single-process result."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "levels.npz")
    mp.spawn(_level_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    fn, costs = _oracle_level_fn()
    ref_hits, ref_stats = fn(list(range(len(costs))))
    assert ref_hits.size > 0 and np.array_equal(got["hits"], sharding.normalise_hits(ref_hits))
    assert tuple(got["stats"]) == tuple(ref_stats)


def _pipelined_worker(rank, world, port, out_path, big=False):
    import torch.distributed as dist
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "oracle"), os.path.join(root, "tests")):
        sys.path.insert(0, p)
    from waldboost_b200._native import HIT_DTYPE
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = sharding.HitGatherer(presorted=True)
    futs = []
    for step in range(4):                               # ragged and empty lists, several steps in flight
        n = (3 * step + 2 * rank) % 5 + (3000 if big and step == 2 and rank == 1 else 0)
        h = np.zeros(n, HIT_DTYPE)
        h["frame"], h["r"], h["c"], h["score"] = 10 * rank + step, np.arange(n), step, rank + 0.5
        futs.append(g.submit(h, (100 * step + rank, 7)))
    res = [f.result() for f in futs]
    g.close()
    if rank == 0:
        np.savez(out_path, **{f"hits{i}": r[0] for i, r in enumerate(res)}, stats=np.array([r[1] for r in res]))
    else:
        assert all(r == (None, None) for r in res)
    dist.destroy_process_group()


@pytest.mark.parametrize("big", [False, True])
def test_pipelined_gather_world2_gloo(tmp_path, big):
    """HitGatherer: gathers submitted back to back come out per step, in rank order, with the counters summed; with `big`
    one rank's list is three orders of magnitude longer than the others' in one step (padding to the longest list)."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "pipe.npz")
    mp.spawn(_pipelined_worker, args=(2, _free_port(), out, big), nprocs=2, join=True)
    got = np.load(out)
    for step in range(4):
        h = got[f"hits{step}"]
        n0, n1 = (3 * step) % 5, (3 * step + 2) % 5 + (3000 if big and step == 2 else 0)
        assert h.size == n0 + n1
        assert np.array_equal(h["frame"], [step] * n0 + [10 + step] * n1) and np.array_equal(h["r"], list(range(n0)) + list(range(n1)))
        assert tuple(got["stats"][step]) == (200 * step + 1, 14)
