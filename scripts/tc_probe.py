"""Times sift_tc_kernel variants (development aid): python scripts/tc_probe.py [images] [frames]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pymcl_b200
from pymcl_b200 import _lib
import bench

ctx = _lib.Context(0)
n_images, n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 200, int(sys.argv[2]) if len(sys.argv) > 2 else 100
modes = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1, 2]
algo = int(sys.argv[4]) if len(sys.argv) > 4 else _lib.ALGO_TENSOR
map_des, frames = bench.make_workload(0, n_images, n_frames)
resident = _lib.MapIndex(ctx, "SIFT", map_des)
q = _lib.QueryBatch(ctx, "SIFT", frames)
import torch
out = torch.zeros((n_frames, n_images), dtype=torch.int32, device="cuda")
ops = 2.0 * 2048 * 2048 * 128 * n_images * n_frames
tiles_per_sm = (n_frames * 2048 / 256) * (n_images * 2048 / 128) / 148  # in 256x128 units: 512 cycles at the int8 peak
for mode in modes:
    ctx.set_option(0, mode)
    ctx.profile_enable(True)
    for _ in range(3):
        resident.match_counts_device(q, out.data_ptr(), algo=algo)
    ctx.sync()
    ms, n = ctx.profile_read("sift_tc")
    ctx.profile_enable(False)
    ms /= n
    print("mode %d: %.3f ms/launch  %.1f TOP/s  %.1f%% of 4500  %.0f cycles/tile-step @1.965GHz" % (
        mode, ms, ops / ms / 1e9, ops / ms / 1e9 / 45.0, ms * 1e-3 * 1.965e9 / tiles_per_sm))
