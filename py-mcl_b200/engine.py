"""Sharding of the map across the GPUs of one box (one process per GPU, torch.distributed).

Every (query descriptor, map image) top-2 is independent of every other map image, so whole images are
partitioned across ranks (contiguous slices balanced by descriptor count); the query batch is identical on
every rank (same files, same deterministic cv2 extraction, or broadcast by the caller); each rank matches
against its slice and ONE all-gather of the int32 per-image counts follows.  There is no
other data-path collective.
"""
from __future__ import annotations

import os
import sys

import numpy as np

from . import _lib


def shard_bounds(rows_per_image, world):
    """Contiguous image ranges [(lo, hi)] * world, balanced by descriptor rows (+1 per image so that empty
    images spread too): image i goes to the rank whose share of the total weight holds i's midpoint.
    Deterministic; every rank computes the same table."""
    n = len(rows_per_image)
    if n == 0:
        return [(0, 0)] * world
    w = np.asarray(rows_per_image, dtype=np.int64) + 1
    csum = np.cumsum(w)
    total = int(csum[-1])
    mid2 = 2 * csum - w                      # 2 * midpoint, exact integers
    owner = np.minimum((mid2 * world) // (2 * total), world - 1)
    edges = np.searchsorted(owner, np.arange(world + 1), side="left")
    return [(int(edges[r]), int(edges[r + 1])) for r in range(world)]


def _dist():
    # a process that never imported torch cannot have initialised torch.distributed: do not pay the import (seconds) for it
    if "torch" not in sys.modules and "WORLD_SIZE" not in os.environ:
        return None
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:
        pass
    return None


def sharded_apply(fn, items, group=None):
    """[fn(x) for x in items], with the calls dealt round-robin to the ranks and the results all-gathered in order.
    Used for cv2 extraction of the query frames, which is deterministic and the end-to-end bottleneck once matching runs
    on the GPU: N ranks extract N times faster and every rank ends up with every frame's descriptors."""
    d = _dist()
    if d is None or d.get_world_size(group) == 1:
        return [fn(x) for x in items]
    world, rank = d.get_world_size(group), d.get_rank(group)
    mine = [fn(x) for x in items[rank::world]]
    parts = [None] * world
    d.all_gather_object(parts, mine, group=group)
    out = [None] * len(items)
    for r in range(world):
        out[r::world] = parts[r]
    return out


class ShardedMap(object):
    """All map images of all locations, flattened, with this rank's slice resident on its GPU."""

    def __init__(self, kind, descs, ctx=None, local_factory=None, group=None):
        self.kind = kind
        self.n_images = len(descs)
        self.group = group
        d = _dist()
        self.world = d.get_world_size(group) if d else 1
        self.rank = d.get_rank(group) if d else 0
        rows = [0 if x is None else len(x) for x in descs]
        self.bounds = shard_bounds(rows, self.world)
        lo, hi = self.bounds[self.rank]
        self.lo, self.hi = lo, hi
        if local_factory is None:
            ctx = ctx or _lib.default_context()
            self.local = _lib.MapIndex(ctx, kind, descs[lo:hi])
        else:
            self.local = local_factory(descs[lo:hi])

    def close(self):
        if hasattr(self.local, "close"):
            self.local.close()

    def match_counts(self, q_descs, mode=_lib.KNN2_RATIO, ratio=0.7, mask=None, fill_value=1, algo=_lib.ALGO_AUTO):
        """(n_queries, n_images) int32 on every rank."""
        lo, hi = self.lo, self.hi
        local_mask = None if mask is None else np.ascontiguousarray(np.asarray(mask)[:, lo:hi])
        local = self.local.match_counts(q_descs, mode=mode, ratio=ratio, mask=local_mask, fill_value=fill_value, algo=algo)
        return self.gather(local)

    def gather(self, local):
        """all-gather of the per-image counts: local (F, hi-lo) -> (F, n_images)."""
        return gather_columns(local, self.bounds, self.group)


_pinned_cache = {}


def _pinned(shape, dtype):
    """A reusable pinned host tensor (NCCL path): pageable .cuda() / .cpu() copies of a few MB cost more than the
    all-gather itself."""
    import torch
    key = (tuple(shape), dtype)
    t = _pinned_cache.get(key)
    if t is None:
        if len(_pinned_cache) > 16:
            _pinned_cache.clear()
        t = _pinned_cache[key] = torch.empty(shape, dtype=dtype).pin_memory()
    return t


def gather_columns(local, bounds, group=None):
    """The one data-path collective: every rank holds columns [lo, hi) = bounds[rank] of an (F, n) result; one
    all_gather_into_tensor of the zero-padded slices gives every rank the full matrix (NCCL on the GPU, gloo in tests)."""
    d = _dist()
    world = d.get_world_size(group) if d else 1
    if world == 1:
        return local
    import torch
    F = local.shape[0]
    maxw = max(h - l for l, h in bounds)
    nccl = d.get_backend(group) == "nccl"
    tdt = torch.from_numpy(np.empty(0, dtype=local.dtype)).dtype
    if nccl:
        stage = _pinned((F, maxw), tdt)
        sv = stage.numpy()
        sv[:, :local.shape[1]] = local
        sv[:, local.shape[1]:] = 0
        t = stage.cuda(non_blocking=True)
    else:
        pad = np.zeros((F, maxw), dtype=local.dtype)
        pad[:, :local.shape[1]] = local
        t = torch.from_numpy(pad)
    out = torch.empty((world * F, maxw), dtype=t.dtype, device=t.device)   # concatenation along dim 0
    d.all_gather_into_tensor(out, t, group=group)
    if nccl:
        host = _pinned((world * F, maxw), tdt)
        host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        out = host.numpy().reshape(world, F, maxw)
    else:
        out = out.numpy().reshape(world, F, maxw)
    full = np.empty((F, bounds[-1][1]), dtype=local.dtype)
    for r, (l, h) in enumerate(bounds):
        full[:, l:h] = out[r][:, :h - l]
    return full


class ShardedBow(object):
    """Bag-of-words indices of all locations, with this rank's LOCATIONS resident on its GPU (a
    location owns its code book + tf-idf index, so locations are the unit; the float32 scores are all-gathered).
    `loader(i)` returns location i's (im_features, idf, voc) -- joblib.load of its .pkl -- and is only called for the
    locations of this rank, so index loading is sharded too."""

    def __init__(self, n_locations, loader, num_words=None, ctx=None, local_factory=None, group=None):
        d = _dist()
        self.group = group
        self.world = d.get_world_size(group) if d else 1
        self.rank = d.get_rank(group) if d else 0
        self.n_loc = int(n_locations)
        self.loc_bounds = shard_bounds([1] * self.n_loc, self.world)
        lo, hi = self.loc_bounds[self.rank]
        mine = [loader(i) for i in range(lo, hi)]
        counts = [int(np.asarray(m[0]).shape[0]) for m in mine]
        widths = [int(np.asarray(m[0]).shape[1]) for m in mine]
        if self.world > 1:
            parts = [None] * self.world
            d.all_gather_object(parts, (counts, widths), group=group)
            counts = [c for part in parts for c in part[0]]
            widths = [w for part in parts for w in part[1]]
        if num_words is None:
            num_words = widths[0] if widths else 0
        self.W = int(num_words)
        self.img_off = np.zeros(self.n_loc + 1, dtype=np.int64)
        np.cumsum(counts, out=self.img_off[1:])
        self.total_images = int(self.img_off[-1])
        # column range of every rank in the (frames, total_images) score matrix
        self.bounds = [(int(self.img_off[l]), int(self.img_off[h])) for l, h in self.loc_bounds]
        if not mine:
            self.local = None
        elif local_factory is None:
            self.local = _lib.BowIndex(ctx or _lib.default_context(), mine, self.W)
        else:
            self.local = local_factory(mine, self.W)

    def close(self):
        if self.local is not None and hasattr(self.local, "close"):
            self.local.close()

    def score(self, q_descs):
        """(n_queries, total_images) float32 on every rank: BOWMatch of every query against every location.
        (Keeping the scores on the device for the all-gather was measured: no faster than this host hand-over at 2 GPUs.)"""
        lo, hi = self.bounds[self.rank]
        if self.local is None:
            local = np.zeros((len(q_descs), 0), dtype=np.float32)
        else:
            local = self.local.score(q_descs)[0]
        assert local.shape[1] == hi - lo
        return gather_columns(local, self.bounds, self.group)
