"""ctypes binding of libmcl.so (the C-ABI in include/mcl.h).  Thin on purpose: numpy in, numpy out.

There is NO CPU fallback: if the shared library is not built, or no sm_100 device is present, every
entry point raises ``MclError``.  Nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmcl.so")

SIFT, ORB, SURF, BOWQ = 0, 1, 2, 3
U8, F32 = 0, 1
KNN2_RATIO, CROSSCHECK = 0, 1
ALGO_AUTO, ALGO_TENSOR, ALGO_SIMT = 0, 1, 2
F_COMMAND, F_PREV, F_BLUR, F_BEST, F_ALL = 1, 2, 4, 8, 15
T_PY, T_F64, T_F32 = 0, 1, 2
KIND_BY_NAME = {"SIFT": SIFT, "ORB": ORB, "SURF": SURF}
ROW_DIM = {SIFT: 128, ORB: 32, SURF: 64, BOWQ: 128}
PROFILE_SLOTS = {"sift_tc": 0, "sift_fixup": 1, "orb": 2, "surf": 3, "bow_vq": 4, "bow_score": 5, "laplacian": 6,
                 "ingest": 7, "sift_simt": 8, "bow_vq_tc": 9, "bow_resolve": 10, "chi2": 11, "orb_fixup": 12, "bow_rescan": 13}


class MclError(RuntimeError):
    pass


_lib = None
_lib_lock = threading.Lock()

_p = C.c_void_p
_i64p = C.POINTER(C.c_int64)

# name -> (restype, argtypes); must list every symbol include/mcl.h declares (tests check this)
SIGNATURES = {
    "mcl_init": (C.c_int, [C.c_int, C.POINTER(_p)]),
    "mcl_shutdown": (None, [_p]),
    "mcl_last_error": (C.c_char_p, []),
    "mcl_device_info": (C.c_int, [_p, _i64p]),
    "mcl_sync": (C.c_int, [_p]),
    "mcl_stats": (C.c_int, [_p, _i64p]),
    "mcl_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(_p)]),
    "mcl_host_free": (None, [_p]),
    "mcl_launch_count": (C.c_int64, [_p]),
    "mcl_set_option": (C.c_int, [_p, C.c_int, C.c_int]),
    "mcl_profile_enable": (C.c_int, [_p, C.c_int]),
    "mcl_profile_read": (C.c_int, [_p, C.c_int, C.POINTER(C.c_double), _i64p]),
    "mcl_map_create": (C.c_int, [_p, C.c_int, C.c_int, _i64p, _p, C.c_int, C.POINTER(_p)]),
    "mcl_map_destroy": (None, [_p]),
    "mcl_map_info": (C.c_int, [_p, _i64p]),
    "mcl_queries_create": (C.c_int, [_p, C.c_int, C.c_int, _i64p, _p, C.c_int, C.POINTER(_p)]),
    "mcl_queries_create_device": (C.c_int, [_p, C.c_int, C.c_int, _i64p, _p, C.c_int, _p, C.POINTER(_p)]),
    "mcl_queries_validate": (C.c_int, [_p]),
    "mcl_queries_destroy": (None, [_p]),
    "mcl_match_counts": (C.c_int, [_p, C.c_int, _i64p, _p, C.c_int, C.c_int, C.c_double, _p, C.c_int32, C.c_int, _p]),
    "mcl_match_counts_device": (C.c_int, [_p, _p, C.c_int, C.c_double, _p, C.c_int32, C.c_int, _p, _p]),
    "mcl_bow_create": (C.c_int, [_p, C.c_int, _i64p, _p, _p, _i64p, _p, C.c_int, C.POINTER(_p)]),
    "mcl_bow_destroy": (None, [_p]),
    "mcl_bow_score": (C.c_int, [_p, C.c_int, _i64p, _p, C.c_int, _p, _p]),
    "mcl_bow_score_device": (C.c_int, [_p, _p, _p, _p, _p]),
    "mcl_kmeans": (C.c_int, [_p, C.c_int64, _p, C.c_int, C.c_int, _p, C.c_double, C.c_int, _p, C.POINTER(C.c_int),
                            C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "mcl_laplacian_var": (C.c_int, [_p, C.c_int, C.POINTER(_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), _p]),
    "mcl_laplacian_var_packed": (C.c_int, [_p, C.c_int, _p, C.POINTER(C.c_int), C.POINTER(C.c_int), _p]),
    "mcl_laplacian_sums": (C.c_int, [_p, C.c_int, C.POINTER(_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), _p]),
    "mcl_filter_step": (C.c_int, [_p, C.c_int, C.c_int, C.c_int, C.c_char, C.c_double, _p, _p, _p, _p]),
    "mcl_filter_step_typed": (C.c_int, [_p, C.c_int, C.c_int, C.c_int, C.c_char, C.c_double, C.c_int, _p, _p, _p, _p, _p, _p, _p]),
    "mcl_chi2": (C.c_int, [_p, C.c_int, C.c_int, C.c_int, _p, _p, C.c_float, _p]),
    "mcl_microbench": (C.c_int, [_p, C.c_int, C.c_int, _p]),
}


def load_library():
    """dlopen libmcl.so and type every entry point.  Does not touch the GPU."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise MclError(
                "libmcl.so is not built (%s).  Build it with `python __graft_entry__.py` "
                "(nvcc -gencode arch=compute_100a,code=sm_100a).  There is no CPU fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def _check(rc, what):
    if rc != 0:
        msg = load_library().mcl_last_error()
        raise MclError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def _offsets(lengths):
    off = np.zeros(len(lengths) + 1, dtype=np.int64)
    if len(lengths):
        np.cumsum(np.asarray(lengths, dtype=np.int64), out=off[1:])
    return off


def _ptr(a):
    return a.ctypes.data_as(_p) if a is not None else None


def pack_descriptors(kind, descs):
    """List of per-image/per-frame descriptor matrices (None or empty allowed, as cv2 returns None for an
    image without keypoints) -> (row offsets int64, concatenated C-contiguous matrix, dtype code)."""
    dim = ROW_DIM[kind]
    mats = []
    for d in descs:
        if d is None or len(d) == 0:
            mats.append(None)
            continue
        d = np.asarray(d)
        if d.ndim != 2 or d.shape[1] != dim:
            raise MclError("descriptor matrix has shape %r, expected (*, %d)" % (d.shape, dim))
        mats.append(d)
    off = _offsets([0 if m is None else len(m) for m in mats])
    real = [m for m in mats if m is not None]
    if kind == ORB:
        dt, code = np.uint8, U8
    elif kind == SURF:
        dt, code = np.float32, F32
    else:  # SIFT / BOWQ: cv2 gives integer-valued float32; uint8 accepted as is
        if real and all(m.dtype == np.uint8 for m in real):
            dt, code = np.uint8, U8
        else:
            dt, code = np.float32, F32
    if real:
        if len(real) == 1 and real[0].dtype == dt and real[0].flags.c_contiguous:
            cat = real[0]
        else:
            cat = _concatenate_staged(real, dt, dim)
    else:
        cat = np.zeros((0, dim), dtype=dt)
    return off, cat, code


_STAGE_MIN_BYTES = 1 << 20
_stage_local = threading.local()


def _concatenate_staged(mats, dt, dim):
    """np.concatenate(mats) -- into a reusable PINNED host buffer when the result is large: the H2D copy that follows is
    then one direct DMA instead of the driver's own staged copy of pageable memory (for a 200-frame batch of small SIFT
    frames the two host copies were most of the end-to-end time), and big batches are copied by a few threads (numpy
    releases the GIL in the copy).  The buffer is per thread and only valid until the next call on that thread; every
    libmcl entry point that consumes it is synchronous.  Without a CUDA runtime (CPU-only tests) it is ordinary memory."""
    rows = sum(len(m) for m in mats)
    nbytes = rows * dim * np.dtype(dt).itemsize
    if nbytes < _STAGE_MIN_BYTES:
        return np.ascontiguousarray(np.concatenate([m.astype(dt, copy=False) for m in mats], axis=0))
    buf = getattr(_stage_local, "buf", None)
    if buf is None or buf.nbytes < nbytes:
        cap = int(nbytes * 1.25) + 4096
        try:
            buf = pinned_empty((cap,), np.uint8)
        except Exception:
            buf = np.empty((cap,), dtype=np.uint8)
        _stage_local.buf = buf
    cat = buf[:nbytes].view(dt).reshape(rows, dim)
    pos, jobs = 0, []
    for m in mats:
        jobs.append((pos, m))
        pos += len(m)

    def copy(group):
        for at, m in group:
            np.copyto(cat[at:at + len(m)], m, casting="unsafe")

    workers = 4
    if nbytes >= (8 << 20) and len(jobs) >= workers:
        from concurrent.futures import ThreadPoolExecutor
        pool = getattr(_stage_local, "pool", None)
        if pool is None:
            pool = _stage_local.pool = ThreadPoolExecutor(max_workers=workers)
        per = (len(jobs) + workers - 1) // workers      # one future per worker, not per matrix
        list(pool.map(copy, [jobs[i:i + per] for i in range(0, len(jobs), per)]))
    else:
        copy(jobs)
    return cat


class Context:
    """One CUDA context/stream on one B200 (mcl_init).  One per process (one process per GPU)."""

    def __init__(self, device=None):
        lib = load_library()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        h = _p()
        _check(lib.mcl_init(int(device), C.byref(h)), "mcl_init")
        self._h = h
        self.device = int(device)
        self._lib = lib
        # development aid: MCL_OPTIONS="4=1,1=0" applies mcl_set_option(which, value) pairs to every new context
        for item in filter(None, os.environ.get("MCL_OPTIONS", "").split(",")):
            which, value = item.split("=")
            self.set_option(int(which), int(value))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mcl_shutdown(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        a = (C.c_int64 * 8)()
        _check(self._lib.mcl_device_info(self._h, a), "mcl_device_info")
        return {"sm_count": a[0], "cc": (a[1], a[2]), "hbm_mib": a[3], "sm_clock_khz": a[4], "l2_bytes": a[5],
                "smem_optin": a[6], "sift_rescans": a[7]}

    def sync(self):
        _check(self._lib.mcl_sync(self._h), "mcl_sync")

    def stats(self):
        a = (C.c_int64 * 8)()
        _check(self._lib.mcl_stats(self._h, a), "mcl_stats")
        return {"sift_candidates": a[0], "sift_rescans": a[1], "repaired_pairs": a[2], "bow_rescanned_rows": a[3],
                "orb_candidates": a[4], "bow_resolve_entries": a[6]}

    def set_option(self, option, value):
        _check(self._lib.mcl_set_option(self._h, int(option), int(value)), "mcl_set_option")

    def launch_count(self):
        return int(self._lib.mcl_launch_count(self._h))

    def profile_enable(self, on=True):
        _check(self._lib.mcl_profile_enable(self._h, 1 if on else 0), "mcl_profile_enable")

    def profile_read(self, which):
        ms, n = C.c_double(), C.c_int64()
        _check(self._lib.mcl_profile_read(self._h, PROFILE_SLOTS[which] if isinstance(which, str) else which,
                                          C.byref(ms), C.byref(n)), "mcl_profile_read")
        return ms.value, n.value

    def microbench(self, which, iters):
        out = (C.c_double * 2)()
        _check(self._lib.mcl_microbench(self._h, which, iters, out), "mcl_microbench")
        return out[0], out[1]

    def kmeans(self, obs, guess, thresh=1e-5, max_iter=0):
        """scipy.cluster.vq._kmeans(obs, guess, thresh) on the GPU -> (code_book (k', 128) float32, avg distortion, iterations)."""
        obs = np.asarray(obs)
        if obs.ndim != 2 or obs.shape[1] != 128:
            raise MclError("kmeans: observations must be (n, 128)")
        if obs.dtype == np.uint8:
            o, code = np.ascontiguousarray(obs), U8
        else:
            o, code = np.ascontiguousarray(obs, dtype=np.float32), F32
        g = np.ascontiguousarray(guess, dtype=np.float32)
        if g.ndim != 2 or g.shape[1] != 128:
            raise MclError("kmeans: guess must be (k, 128)")
        out = np.zeros_like(g)
        k_out, dist, iters = C.c_int(), C.c_double(), C.c_int()
        _check(self._lib.mcl_kmeans(self._h, len(o), _ptr(o), code, len(g), _ptr(g), float(thresh), int(max_iter), _ptr(out),
                                    C.byref(k_out), C.byref(dist), C.byref(iters)), "mcl_kmeans")
        return out[:k_out.value].copy(), dist.value, iters.value

    # ---- small stages -------------------------------------------------------------------
    def laplacian_var(self, images):
        """analyzer.Laplacian (analyze.py:441-445) for a list of 2-D uint8 arrays -> float64 array."""
        return self._lap(images, False)

    def laplacian_sums(self, images):
        return self._lap(images, True)

    def _lap(self, images, sums):
        n = len(images)
        imgs = [np.asarray(im, dtype=np.uint8) for im in images]
        for im in imgs:
            if im.ndim != 2:
                raise MclError("laplacian: images must be 2-D grayscale uint8")
        total = sum(im.shape[0] * ((im.shape[1] + 15) & ~15) for im in imgs)
        if not sums and n > 1 and total >= _STAGE_MIN_BYTES:
            return self._lap_packed(imgs, total)
        imgs = [np.ascontiguousarray(im) for im in imgs]
        ptrs = (_p * max(n, 1))(*[im.ctypes.data for im in imgs])
        hs = (C.c_int * max(n, 1))(*[im.shape[0] for im in imgs])
        ws = (C.c_int * max(n, 1))(*[im.shape[1] for im in imgs])
        st = (C.c_int * max(n, 1))(*[im.strides[0] for im in imgs])
        if sums:
            out = np.zeros((n, 2), dtype=np.int64)
            _check(self._lib.mcl_laplacian_sums(self._h, n, ptrs, hs, ws, st, _ptr(out)), "mcl_laplacian_sums")
        else:
            out = np.zeros(n, dtype=np.float64)
            _check(self._lib.mcl_laplacian_var(self._h, n, ptrs, hs, ws, st, _ptr(out)), "mcl_laplacian_var")
        return out

    def _lap_packed(self, imgs, total):
        """A sequence's worth of frames: the images are copied (a few threads; numpy releases the GIL) into a reusable PINNED
        buffer in the device's own layout -- rows padded to 16 bytes, images back to back -- and go to the GPU in one DMA
        instead of one pageable strided copy per image (which bounded the call at 7.4 GB/s)."""
        n = len(imgs)
        buf = getattr(_stage_local, "lap", None)
        if buf is None or buf.nbytes < total:
            try:
                buf = pinned_empty((int(total * 1.25) + 4096,), np.uint8)
            except Exception:
                buf = np.empty((int(total * 1.25) + 4096,), dtype=np.uint8)
            _stage_local.lap = buf
        jobs, pos = [], 0
        for im in imgs:
            h, w = im.shape
            pitch = (w + 15) & ~15
            jobs.append((pos, h, w, pitch, im))
            pos += h * pitch

        def copy(group):
            for at, h, w, pitch, im in group:
                np.copyto(buf[at:at + h * pitch].reshape(h, pitch)[:, :w], im)

        workers = 4
        if n >= workers and total >= (8 << 20):
            from concurrent.futures import ThreadPoolExecutor
            pool = getattr(_stage_local, "pool", None)
            if pool is None:
                pool = _stage_local.pool = ThreadPoolExecutor(max_workers=workers)
            per = (n + workers - 1) // workers
            list(pool.map(copy, [jobs[i:i + per] for i in range(0, n, per)]))
        else:
            copy(jobs)
        hs = (C.c_int * n)(*[im.shape[0] for im in imgs])
        ws = (C.c_int * n)(*[im.shape[1] for im in imgs])
        out = np.zeros(n, dtype=np.float64)
        _check(self._lib.mcl_laplacian_var_packed(self._h, n, _ptr(buf), hs, ws, _ptr(out)), "mcl_laplacian_var_packed")
        return out

    def filter_step(self, stages, cmd, blur, prev, cur, angle_step=15):
        """prev, cur: (L, 1+I) float64.  Returns (prev after the command shift, out, (bestCircle, bestAngle))."""
        self.set_option(6, int(angle_step))
        prev = np.ascontiguousarray(prev, dtype=np.float64).copy()
        cur = np.ascontiguousarray(cur, dtype=np.float64)
        if prev.shape != cur.shape or prev.ndim != 2:
            raise MclError("filter_step: prev/cur must both be (L, 1+I)")
        out = np.zeros_like(prev)
        best = np.zeros(2, dtype=np.int32)
        c = (cmd or "s").encode()[:1]
        _check(self._lib.mcl_filter_step(self._h, prev.shape[0], prev.shape[1] - 1, int(stages), c, float(blur),
                                         _ptr(prev), _ptr(cur), _ptr(out), _ptr(best)), "mcl_filter_step")
        return prev, out, (int(best[0]), int(best[1]))

    def filter_step_typed(self, stages, cmd, blur, blur_type, prev, prev_type, cur, cur_type, angle_step=15):
        """filter_step on lists that hold numpy scalars (include/mcl.h: mcl_filter_step_typed).  *_type: (L, 2) int8 tags
        [total, probabilities] per row (T_PY / T_F64 / T_F32).  Returns (prev, out, out_type, best)."""
        self.set_option(6, int(angle_step))
        prev = np.ascontiguousarray(prev, dtype=np.float64).copy()
        cur = np.ascontiguousarray(cur, dtype=np.float64)
        if prev.shape != cur.shape or prev.ndim != 2:
            raise MclError("filter_step: prev/cur must both be (L, 1+I)")
        pt = np.ascontiguousarray(prev_type, dtype=np.int8)
        ct = np.ascontiguousarray(cur_type, dtype=np.int8)
        if pt.shape != (prev.shape[0], 2) or ct.shape != pt.shape:
            raise MclError("filter_step_typed: type tags must be (L, 2)")
        out = np.zeros_like(prev)
        ot = np.zeros_like(pt)
        best = np.zeros(2, dtype=np.int32)
        c = (cmd or "s").encode()[:1]
        _check(self._lib.mcl_filter_step_typed(self._h, prev.shape[0], prev.shape[1] - 1, int(stages), c, float(blur),
                                               int(blur_type), _ptr(prev), _ptr(pt), _ptr(cur), _ptr(ct), _ptr(out), _ptr(ot),
                                               _ptr(best)), "mcl_filter_step_typed")
        return prev, out, ot, (int(best[0]), int(best[1]))

    def chi2(self, q_hist, map_hist, eps=1e-10):
        """(n_q, bins) x (n_map, bins) float32 histograms -> (n_q, n_map) float32 chi-squared distances, bit-identical to
        the reference's Searcher.chisquared (search.py:22-26) on float32 histograms."""
        q = np.ascontiguousarray(q_hist, dtype=np.float32)
        m = np.ascontiguousarray(map_hist, dtype=np.float32)
        if q.ndim != 2 or m.ndim != 2 or q.shape[1] != m.shape[1]:
            raise MclError("chi2: histograms must be (n, bins) with equal bins")
        out = np.zeros((q.shape[0], m.shape[0]), dtype=np.float32)
        _check(self._lib.mcl_chi2(self._h, q.shape[0], m.shape[0], q.shape[1], _ptr(q), _ptr(m), float(np.float32(eps)),
                                  _ptr(out)), "mcl_chi2")
        return out


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


class MapIndex:
    """Resident descriptors of a set of map images (mcl_map): replaces the {path: (kp, des)} dicts of
    Matcher.createFeatureIndex (Matcher.py:147-163) on the device."""

    def __init__(self, ctx, kind, descs):
        self.ctx = ctx
        self.kind = KIND_BY_NAME[kind] if isinstance(kind, str) else int(kind)
        off, cat, code = pack_descriptors(self.kind, descs)
        self.n_images = len(descs)
        self.row_offsets = off
        h = _p()
        _check(ctx._lib.mcl_map_create(ctx._h, self.kind, self.n_images, off.ctypes.data_as(_i64p), _ptr(cat), code,
                                       C.byref(h)), "mcl_map_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            self.ctx._lib.mcl_map_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        a = (C.c_int64 * 4)()
        _check(self.ctx._lib.mcl_map_info(self._h, a), "mcl_map_info")
        return {"n_images": a[0], "rows": a[1], "padded_rows": a[2], "resident_bytes": a[3]}

    def match_counts(self, q_descs, mode=KNN2_RATIO, ratio=0.7, mask=None, fill_value=1, algo=ALGO_AUTO, packed=None):
        """counts[q, i] = len(good) of *Match(query q, image i); host in / host out (H2D + kernels + D2H).
        `packed` = (offsets, matrix, dtype code) skips the packing of q_descs."""
        off, cat, code = packed if packed is not None else pack_descriptors(self.kind, q_descs)
        nq = len(off) - 1
        out = np.zeros((nq, self.n_images), dtype=np.int32)
        m = None
        if mask is not None:
            m = np.ascontiguousarray(mask, dtype=np.uint8)
            if m.shape != out.shape:
                raise MclError("mask must have shape (n_queries, n_images)")
        _check(self.ctx._lib.mcl_match_counts(self._h, nq, off.ctypes.data_as(_i64p), _ptr(cat), code, int(mode),
                                              float(ratio), _ptr(m), int(fill_value), int(algo), _ptr(out)),
               "mcl_match_counts")
        return out

    def match_counts_device(self, queries, d_counts_ptr, mode=KNN2_RATIO, ratio=0.7, d_mask_ptr=None, fill_value=1,
                            algo=ALGO_AUTO, stream=None):
        """Asynchronous: resident queries, raw device pointers (e.g. torch tensor .data_ptr())."""
        _check(self.ctx._lib.mcl_match_counts_device(self._h, queries._h, int(mode), float(ratio),
                                                     _p(d_mask_ptr) if d_mask_ptr else None, int(fill_value), int(algo),
                                                     _p(d_counts_ptr), _p(stream) if stream else None),
               "mcl_match_counts_device")


class QueryBatch:
    """Resident query descriptors of a batch of frames (mcl_queries)."""

    def __init__(self, ctx, kind, descs=None, packed=None, device=None):
        """descs / packed: host descriptors.  device = (row offsets int64 host array, device pointer of the concatenated
        rows, dtype code, stream or None): the rows are already in device memory (mcl_queries_create_device; asynchronous)."""
        self.ctx = ctx
        self.kind = (KIND_BY_NAME.get(kind, BOWQ) if isinstance(kind, str) else int(kind))
        h = _p()
        if device is not None:
            off, d_ptr, code, stream = device
            off = np.ascontiguousarray(off, dtype=np.int64)
            _check(ctx._lib.mcl_queries_create_device(ctx._h, self.kind, len(off) - 1, off.ctypes.data_as(_i64p), _p(d_ptr), code,
                                                      _p(stream) if stream else None, C.byref(h)), "mcl_queries_create_device")
        else:
            off, cat, code = packed if packed is not None else pack_descriptors(self.kind, descs)
            _check(ctx._lib.mcl_queries_create(ctx._h, self.kind, len(off) - 1, off.ctypes.data_as(_i64p), _ptr(cat), code,
                                               C.byref(h)), "mcl_queries_create")
        self.n_queries = len(off) - 1
        self.row_offsets = off
        self._h = h

    def validate(self):
        """Device-built batches: raises if the ingest saw a float32 SIFT value that is not an integer in [0, 255]."""
        _check(self.ctx._lib.mcl_queries_validate(self._h), "mcl_queries_validate")

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            self.ctx._lib.mcl_queries_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BowIndex:
    """Resident bag-of-words indices of n locations (mcl_bow): per location the (im_features, idf, voc) of
    joblib.load(indexPath) in Matcher.BOWMatch (Matcher.py:236)."""

    def __init__(self, ctx, locations, num_words):
        """locations: list of (im_features (N,W) f32, idf (W,) f32, voc (<=W,128) f32)."""
        self.ctx = ctx
        self.W = int(num_words)
        self.n_loc = len(locations)
        vocs = [np.ascontiguousarray(v, dtype=np.float32) for (_, _, v) in locations]
        idfs = [np.ascontiguousarray(i, dtype=np.float32).reshape(-1) for (_, i, _) in locations]
        imfs = [np.ascontiguousarray(f, dtype=np.float32) for (f, _, _) in locations]
        for i, f in zip(idfs, imfs):
            if i.shape[0] != self.W or f.shape[1] != self.W:
                raise MclError("idf / im_features width != num_words")
        self.voc_off = _offsets([len(v) for v in vocs])
        self.img_off = _offsets([len(f) for f in imfs])
        self.total_images = int(self.img_off[-1])
        voc = np.ascontiguousarray(np.concatenate(vocs, axis=0))
        idf = np.ascontiguousarray(np.stack(idfs, axis=0))
        imf = np.ascontiguousarray(np.concatenate(imfs, axis=0))
        h = _p()
        _check(ctx._lib.mcl_bow_create(ctx._h, self.n_loc, self.voc_off.ctypes.data_as(_i64p), _ptr(voc), _ptr(idf),
                                       self.img_off.ctypes.data_as(_i64p), _ptr(imf), self.W, C.byref(h)), "mcl_bow_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            self.ctx._lib.mcl_bow_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def score_device(self, queries, d_scores_ptr, d_words_ptr=None, stream=None):
        """Asynchronous: resident queries (QueryBatch of kind BOWQ), raw device pointers (e.g. torch tensor .data_ptr()):
        scores (n_queries, total_images) float32, optional words (n_loc, total rows) int32."""
        _check(self.ctx._lib.mcl_bow_score_device(self._h, queries._h, _p(d_words_ptr) if d_words_ptr else None,
                                                  _p(d_scores_ptr), _p(stream) if stream else None), "mcl_bow_score_device")

    def score(self, q_descs, want_words=False, packed=None):
        """-> (scores (n_queries, total_images) f32, words (n_loc, total rows) int32 or None).  `packed` = (offsets, matrix,
        dtype code) skips the packing (e.g. rows already in one pinned buffer)."""
        off, cat, code = packed if packed is not None else pack_descriptors(BOWQ, q_descs)
        nq = len(off) - 1
        scores = np.zeros((nq, self.total_images), dtype=np.float32)
        words = np.zeros((self.n_loc, int(off[-1])), dtype=np.int32) if want_words else None
        _check(self.ctx._lib.mcl_bow_score(self._h, nq, off.ctypes.data_as(_i64p), _ptr(cat), code, _ptr(words),
                                           _ptr(scores)), "mcl_bow_score")
        return scores, words


def pinned_empty(shape, dtype):
    """numpy array over cudaHostAlloc'd memory (for the end-to-end path).  Keeps the allocation alive through
    the array's base object."""
    lib = load_library()
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = _p()
    _check(lib.mcl_host_alloc(max(n, 1), C.byref(p)), "mcl_host_alloc")

    class _Owner:
        def __init__(self, ptr):
            self.ptr = ptr

        def __del__(self):
            try:
                lib.mcl_host_free(self.ptr)
            except Exception:
                pass

    owner = _Owner(p)
    buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
    buf._owner = owner
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    return arr
