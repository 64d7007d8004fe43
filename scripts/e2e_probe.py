"""Times the host-buffer path and its kernels (development aid): python scripts/e2e_probe.py [images] [frames] [algo]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pymcl_b200
from pymcl_b200 import _lib
import bench

ctx = _lib.Context(0)
n_images = int(sys.argv[1]) if len(sys.argv) > 1 else 400
n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 200
algo = int(sys.argv[3]) if len(sys.argv) > 3 else 1
map_des, frames = bench.make_workload(0, n_images, n_frames)
resident = _lib.MapIndex(ctx, "SIFT", map_des)
packed = _lib.pack_descriptors(_lib.SIFT, frames)
pin = _lib.pinned_empty(packed[1].shape, np.float32); pin[...] = packed[1]
pp = (packed[0], pin, packed[2])
q = _lib.QueryBatch(ctx, "SIFT", packed=pp)
import torch
out = torch.zeros((n_frames, n_images), dtype=torch.int32, device="cuda")
for it in range(3):
    ctx.profile_enable(True)
    t0 = time.perf_counter()
    c = resident.match_counts(None, algo=algo, packed=pp)
    t1 = time.perf_counter()
    parts = {k: ctx.profile_read(k) for k in ("sift_tc", "sift_fixup", "sift_simt", "ingest")}
    t2 = time.perf_counter()
    resident.match_counts_device(q, out.data_ptr(), algo=algo)
    ctx.sync()
    t3 = time.perf_counter()
    parts2 = {k: ctx.profile_read(k) for k in ("sift_tc", "sift_fixup", "sift_simt")}
    ctx.profile_enable(False)
    print("host path %.1f ms %s | resident %.1f ms %s | equal %s" % (
        1e3 * (t1 - t0), {k: (round(v[0], 2), v[1]) for k, v in parts.items()}, 1e3 * (t3 - t2),
        {k: (round(v[0], 2), v[1]) for k, v in parts2.items()}, bool((c == out.cpu().numpy()).all())))
