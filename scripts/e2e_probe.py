"""Times the pieces of the host-buffer path (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pymcl_b200
from pymcl_b200 import _lib, synth
import bench

ctx = _lib.Context(0)
n_images, n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 400, int(sys.argv[2]) if len(sys.argv) > 2 else 200
map_des, frames = bench.make_workload(0, n_images, n_frames)
resident = _lib.MapIndex(ctx, "SIFT", map_des)
packed = _lib.pack_descriptors(_lib.SIFT, frames)
pin = _lib.pinned_empty(packed[1].shape, np.float32); pin[...] = packed[1]
pp = (packed[0], pin, packed[2])
for it in range(4):
    t0 = time.perf_counter()
    q = _lib.QueryBatch(ctx, "SIFT", packed=pp)
    t1 = time.perf_counter()
    q.close()
    t2 = time.perf_counter()
    c = resident.match_counts(None, algo=_lib.ALGO_TENSOR, packed=pp)
    t3 = time.perf_counter()
    print("create %.1f ms  destroy %.1f ms  match_counts(host) %.1f ms" % (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2)))
