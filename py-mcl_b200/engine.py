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


def _stream_handle(stream):
    """cudaStream_t of a torch stream for libmcl's *_device calls.  torch's default stream has handle 0, which those calls read
    as "the context's own stream" (not ordered with torch's work): cudaStreamLegacy (0x1) names the default stream explicitly."""
    return stream.cuda_stream or 1


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
    """All map images of all locations, flattened, with this rank's slice resident on its GPU.

    One process per GPU.  With NCCL the whole call stays on the device: every query frame is uploaded by ONE rank and its
    rows are all-gathered over NVLink (instead of every rank pulling the same batch through its own PCIe link), each rank
    matches against its slice (mcl_match_counts_device), the int32 counts are all-gathered on the device, and ONE
    device-to-host copy returns them.  Other backends (gloo in the CPU tests, or a custom local matcher) take the host path."""

    # batches below this many bytes are uploaded whole by every rank (the collective would cost more than the copy)
    SHARED_UPLOAD_MIN_BYTES = 8 << 20

    def __init__(self, kind, descs, ctx=None, local_factory=None, group=None, local_only=False):
        """descs: every map image's descriptors (each rank keeps its slice), or -- local_only=True -- just THIS rank's
        images, in rank order (the slices' sizes are exchanged once)."""
        self.kind = kind
        self.group = group
        d = _dist()
        self.world = d.get_world_size(group) if d else 1
        self.rank = d.get_rank(group) if d else 0
        if local_only:
            counts = [len(descs)]
            if self.world > 1:
                counts = [None] * self.world
                d.all_gather_object(counts, len(descs), group=group)
            edges = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
            self.bounds = [(int(edges[r]), int(edges[r + 1])) for r in range(self.world)]
            self.n_images = int(edges[-1])
            local = descs
        else:
            self.n_images = len(descs)
            rows = [0 if x is None else len(x) for x in descs]
            self.bounds = shard_bounds(rows, self.world)
            lo, hi = self.bounds[self.rank]
            local = descs[lo:hi]
        self.lo, self.hi = self.bounds[self.rank]
        if local_factory is None:
            ctx = ctx or _lib.default_context()
            self.local = _lib.MapIndex(ctx, kind, local)
        else:
            self.local = local_factory(local)
        self._ctx = ctx
        self._bufs = {}

    def close(self):
        if hasattr(self.local, "close"):
            self.local.close()
        self._bufs = {}

    # ---- host path (single process, gloo, custom local matcher) ---------------------------------------------------
    def match_counts(self, q_descs, mode=_lib.KNN2_RATIO, ratio=0.7, mask=None, fill_value=1, algo=_lib.ALGO_AUTO, packed=None):
        """(n_queries, n_images) int32 on every rank.  `packed` = (offsets, matrix, dtype code) skips the packing."""
        if self._on_device():
            return self._match_counts_nccl(q_descs, mode, ratio, mask, fill_value, algo, packed)
        lo, hi = self.lo, self.hi
        local_mask = None if mask is None else np.ascontiguousarray(np.asarray(mask)[:, lo:hi])
        kw = {} if packed is None else {"packed": packed}
        local = self.local.match_counts(q_descs, mode=mode, ratio=ratio, mask=local_mask, fill_value=fill_value, algo=algo, **kw)
        return self.gather(local)

    def gather(self, local):
        """all-gather of the per-image counts: local (F, hi-lo) -> (F, n_images)."""
        return gather_columns(local, self.bounds, self.group)

    # ---- device path (NCCL) -------------------------------------------------------------------------------------------
    def _on_device(self):
        d = _dist()
        return (self.world > 1 and d is not None and d.get_backend(self.group) == "nccl" and isinstance(self.local, _lib.MapIndex))

    def _buf(self, name, shape, dtype, pinned=False):
        """grow-only cache of device / pinned host tensors (no allocation on the per-call path)."""
        import torch
        n = int(np.prod(shape))
        t = self._bufs.get(name)
        if t is None or t.numel() < n or t.dtype != dtype:
            cap = max(n, 1)
            t = torch.empty(cap, dtype=dtype).pin_memory() if pinned else torch.empty(cap, dtype=dtype, device="cuda")
            self._bufs[name] = t
        return t[:n].view(*shape)

    def match_counts_device(self, queries, mode=_lib.KNN2_RATIO, ratio=0.7, d_mask=None, fill_value=1, algo=_lib.ALGO_AUTO):
        """Resident query batch (a _lib.QueryBatch on this rank's context) -> torch int32 tensor (n_queries, n_images) on the
        device, valid on the current torch stream: this rank's counts + the device all-gather, nothing touches the host."""
        import torch
        d = _dist()
        F, w = queries.n_queries, self.hi - self.lo
        maxw = max(h - l for l, h in self.bounds)
        stream = torch.cuda.current_stream()
        mine = self._buf("counts_local", (F, maxw), torch.int32)
        if w < maxw:
            mine.zero_()
        tight = mine if w == maxw else self._buf("counts_tight", (F, w), torch.int32)
        if w:
            self.local.match_counts_device(queries, tight.data_ptr(), mode=mode, ratio=ratio,
                                           d_mask_ptr=d_mask.data_ptr() if d_mask is not None else None, fill_value=fill_value,
                                           algo=algo, stream=_stream_handle(stream))
            if tight is not mine:
                mine[:, :w].copy_(tight)
        if self.world == 1:
            return mine
        allc = self._buf("counts_all", (self.world, F, maxw), torch.int32)
        d.all_gather_into_tensor(allc, mine, group=self.group)
        if all(h - l == maxw for l, h in self.bounds):
            return allc.permute(1, 0, 2).reshape(F, self.world * maxw)
        full = self._buf("counts_full", (F, self.n_images), torch.int32)
        for r, (l, h) in enumerate(self.bounds):
            full[:, l:h].copy_(allc[r][:, :h - l])
        return full

    def _match_counts_nccl(self, q_descs, mode, ratio, mask, fill_value, algo, packed):
        import torch
        d = _dist()
        off, cat, code = packed if packed is not None else _lib.pack_descriptors(self.local.kind, q_descs)
        F = len(off) - 1
        if F == 0:
            return np.zeros((0, self.n_images), dtype=np.int32)
        rows = int(off[-1])
        row_elems = cat.shape[1] if cat.ndim == 2 else _lib.ROW_DIM[self.local.kind]
        tdt = torch.from_numpy(np.empty(0, dtype=cat.dtype)).dtype
        main = torch.cuda.current_stream()
        tl = [] if os.environ.get("MCL_ENGINE_TIMELINE") else None   # (label, timing event): device timeline of the call

        def mark(label, stream):
            if tl is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(stream)
                tl.append((label, ev))

        mark("start", main)
        d_mask = None
        if mask is not None:
            m = np.ascontiguousarray(np.asarray(mask)[:, self.lo:self.hi], dtype=np.uint8)
            d_mask = self._buf("mask", m.shape, torch.uint8)
            d_mask.copy_(torch.from_numpy(m), non_blocking=False)
        dq = self._buf("q_rows", (rows, row_elems), tdt)
        src = torch.from_numpy(cat) if rows else None
        plan = None
        if cat.nbytes >= self.SHARED_UPLOAD_MIN_BYTES and F >= 2 * self.world:
            # frames in two parts (1/8, 7/8 of the rows): only the first, small part's upload + all-gather is exposed; the
            # second part's runs on a side stream under the first part's matching.  (Three parts, like mcl_match_counts' own
            # host pipeline, measured worse at 4 GPUs: 1.5 instead of 1.0 ms over the resident step -- every part is another
            # set of launches with its own tail.)  Inside a part the frames are dealt to the ranks in contiguous runs of
            # nearly equal row counts; each rank uploads its run, then the runs are all-gathered over NVLink.
            cuts = [0]
            for frac in (1.0 / 8,):
                c = int(np.searchsorted(off, int(rows * frac), side="left"))
                c = min(max(c, cuts[-1] + self.world), F - self.world)
                if c > cuts[-1]:
                    cuts.append(c)
            cuts.append(F)
            plan = []
            for f0, f1 in zip(cuts[:-1], cuts[1:]):
                r0, r1 = int(off[f0]), int(off[f1])
                edges = [r0 + (r1 - r0) * k // self.world for k in range(self.world + 1)]
                fe = [int(np.searchsorted(off, e, side="left")) for e in edges]
                fe[0], fe[-1] = f0, f1
                runs = [(int(off[fe[k]]), int(off[fe[k + 1]])) for k in range(self.world)]
                if any(y <= x for x, y in runs):   # wildly uneven frames: a rank would have nothing to send
                    plan = None
                    break
                plan.append((f0, f1, runs))
        if plan is None:
            if rows:
                dq.copy_(src, non_blocking=True)   # small batch (a DOR step): every rank uploads it whole
            parts = [(0, F)]
            events = [None]
        else:
            parts = [(f0, f1) for f0, f1, _ in plan]
            side = self._bufs.get("side_stream")
            if side is None:
                side = self._bufs["side_stream"] = torch.cuda.Stream()
            side.wait_stream(main)
            events = []
            with torch.cuda.stream(side):
                for pi, (f0, f1, runs) in enumerate(plan):
                    # every rank's run lands in its slot of one (world, longest run, row) buffer: this rank's straight from
                    # the host, the others' by ONE in-place all_gather_into_tensor; the runs are then packed back to back
                    # (device copies) so that the part is one contiguous batch
                    maxrun = max(y - x for x, y in runs)
                    g = self._buf("gather%d" % pi, (self.world, maxrun, row_elems), tdt)
                    a, b = runs[self.rank]
                    g[self.rank, :b - a].copy_(src[a:b], non_blocking=True)
                    mark("part %d own run uploaded" % pi, side)
                    d.all_gather_into_tensor(g.view(self.world * maxrun, row_elems), g[self.rank], group=self.group)
                    for r, (x, y) in enumerate(runs):
                        dq[x:y].copy_(g[r, :y - x], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(side)
                    events.append(ev)
                    mark("part %d rows on every rank (side stream)" % pi, side)
        # this rank's counts of every part land in ONE (F, widest shard) block, gathered once at the end: a gather per part put a
        # rendezvous of all ranks in the middle of the call (at 8 GPUs the step waited there for the slowest process)
        w = self.hi - self.lo
        maxw = max(h - l for l, h in self.bounds)
        mine = self._buf("counts_local", (F, maxw), torch.int32)
        if w < maxw:
            mine.zero_()
        tight = mine if w == maxw else self._buf("counts_tight", (F, max(w, 1)), torch.int32)
        batches = []
        for (f0, f1), ev in zip(parts, events):
            if ev is not None:
                main.wait_event(ev)
            r0 = int(off[f0])
            q = _lib.QueryBatch(self.local.ctx, self.local.kind,
                                device=(off[f0:f1 + 1] - r0, dq.data_ptr() + r0 * row_elems * cat.dtype.itemsize, code, _stream_handle(main)))
            batches.append(q)
            mark("frames %d-%d ingested, matching starts" % (f0, f1), main)
            if w:
                self.local.match_counts_device(q, tight[f0:f1].data_ptr(), mode=mode, ratio=ratio,
                                               d_mask_ptr=None if d_mask is None else d_mask[f0:f1].data_ptr(),
                                               fill_value=fill_value, algo=algo, stream=_stream_handle(main))
            mark("frames %d-%d matched" % (f0, f1), main)
        if w and tight is not mine:
            mine[:, :w].copy_(tight)
        allc = self._buf("counts_all", (self.world, F, maxw), torch.int32)
        d.all_gather_into_tensor(allc, mine, group=self.group)
        counts = self._buf("counts_out", (F, self.n_images), torch.int32)
        if all(h - l == maxw for l, h in self.bounds):
            counts.view(F, self.world, maxw).copy_(allc.permute(1, 0, 2))
        else:
            for r, (l, h) in enumerate(self.bounds):
                if h > l:
                    counts[:, l:h].copy_(allc[r][:, :h - l])
        mark("counts gathered", main)
        host = self._buf("counts_host", (F, self.n_images), torch.int32, pinned=True)
        host.copy_(counts, non_blocking=True)
        mark("counts on the host", main)
        main.synchronize()
        if tl is not None and self.rank == 0:
            import sys
            side_s = self._bufs.get("side_stream")
            if side_s is not None:
                side_s.synchronize()
            print("[engine] " + "; ".join("%s @ %.2f ms" % (lab, tl[0][1].elapsed_time(ev)) for lab, ev in tl[1:]), file=sys.stderr)
        for q in batches:
            q.validate()
            q.close()
        return host.numpy().copy()


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

    def _on_device(self):
        d = _dist()
        return (self.world > 1 and d is not None and d.get_backend(self.group) == "nccl" and
                (self.local is None or isinstance(self.local, _lib.BowIndex)))

    def score(self, q_descs, packed=None):
        """(n_queries, total_images) float32 on every rank: BOWMatch of every query against every location.  With NCCL the
        scores stay on the device: mcl_bow_score_device, one all-gather of the (zero-padded) column blocks, one D2H.
        `packed` = (offsets, matrix, dtype code) skips the packing."""
        lo, hi = self.bounds[self.rank]
        if self._on_device():
            return self._score_nccl(q_descs, packed)
        if self.local is None:
            local = np.zeros((len(packed[0]) - 1 if packed is not None else len(q_descs), 0), dtype=np.float32)
        else:
            local = self.local.score(q_descs, packed=packed)[0] if packed is not None else self.local.score(q_descs)[0]
        assert local.shape[1] == hi - lo
        return gather_columns(local, self.bounds, self.group)

    def score_device(self, queries):
        """Resident query batch (a _lib.QueryBatch of kind BOWQ on this rank's context) -> torch float32 tensor
        (world, n_queries, widest rank's images) on the device, valid on the current torch stream: this rank's kernels + the
        device all-gather (rank r's scores are [r, :, :bounds[r][1] - bounds[r][0]]); nothing touches the host."""
        import torch
        d = _dist()
        F = queries.n_queries
        lo, hi = self.bounds[self.rank]
        w = hi - lo
        maxw = max(h - l for l, h in self.bounds)
        key = (F, maxw)
        if getattr(self, "_key", None) != key:
            self._mine = torch.zeros((F, maxw), dtype=torch.float32, device="cuda")
            self._tight = self._mine if w == maxw else torch.zeros((F, max(w, 1)), dtype=torch.float32, device="cuda")
            self._all = torch.empty((self.world, F, maxw), dtype=torch.float32, device="cuda")
            self._host = torch.empty((self.world, F, maxw), dtype=torch.float32).pin_memory()
            self._key = key
        stream = torch.cuda.current_stream()
        if self.local is not None and w:
            self.local.score_device(queries, self._tight.data_ptr(), stream=_stream_handle(stream))
            if self._tight is not self._mine:
                self._mine[:, :w].copy_(self._tight[:, :w])
        if self.world == 1:
            return self._mine[None]
        d.all_gather_into_tensor(self._all, self._mine, group=self.group)
        return self._all

    def _score_nccl(self, q_descs, packed=None):
        import torch
        d = _dist()
        F = len(packed[0]) - 1 if packed is not None else len(q_descs)
        lo, hi = self.bounds[self.rank]
        w = hi - lo
        maxw = max(h - l for l, h in self.bounds)
        key = (F, maxw)
        if getattr(self, "_key", None) != key:
            self._mine = torch.zeros((F, maxw), dtype=torch.float32, device="cuda")
            self._tight = self._mine if w == maxw else torch.zeros((F, max(w, 1)), dtype=torch.float32, device="cuda")
            self._all = torch.empty((self.world, F, maxw), dtype=torch.float32, device="cuda")
            self._host = torch.empty((self.world, F, maxw), dtype=torch.float32).pin_memory()
            self._key = key
        stream = torch.cuda.current_stream()
        q = None
        if self.local is not None and w:
            # the batch is built from a cached device tensor (pooled blocks inside the library): a host-built batch would
            # cudaMalloc / cudaFree per call, which costs milliseconds once NCCL has mapped its peers
            off, cat, code = packed if packed is not None else _lib.pack_descriptors(_lib.BOWQ, q_descs)
            tdt = torch.from_numpy(np.empty(0, dtype=cat.dtype)).dtype
            if getattr(self, "_dq", None) is None or self._dq.numel() < cat.size or self._dq.dtype != tdt:
                self._dq = torch.empty(max(cat.size, 1), dtype=tdt, device="cuda")
            dq = self._dq[:cat.size]
            if cat.size:
                dq.copy_(torch.from_numpy(cat).reshape(-1), non_blocking=True)
            q = _lib.QueryBatch(self.local.ctx, _lib.BOWQ, device=(off, dq.data_ptr(), code, _stream_handle(stream)))
            self.local.score_device(q, self._tight.data_ptr(), stream=_stream_handle(stream))
            if self._tight is not self._mine:
                self._mine[:, :w].copy_(self._tight[:, :w])
        d.all_gather_into_tensor(self._all, self._mine, group=self.group)
        # the ranks' column blocks are put side by side ON THE DEVICE, then one D2H of the finished (F, total_images) matrix
        if getattr(self, "_full", None) is None or self._full.shape != (F, self.total_images):
            self._full = torch.empty((F, self.total_images), dtype=torch.float32, device="cuda")
            self._full_host = torch.empty((F, self.total_images), dtype=torch.float32).pin_memory()
        if all(h - l == maxw for l, h in self.bounds):   # equal blocks: one strided copy instead of one per rank
            self._full.view(F, self.world, maxw).copy_(self._all.permute(1, 0, 2))
        else:
            for r, (l, h) in enumerate(self.bounds):
                if h > l:
                    self._full[:, l:h].copy_(self._all[r][:, :h - l])
        self._full_host.copy_(self._full, non_blocking=True)
        stream.synchronize()
        if q is not None:
            q.close()
        return self._full_host.numpy().copy()
