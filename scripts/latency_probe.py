"""Single-frame latency of Matcher.run()'s GPU part (1 frame vs 25 images) for the SIFT algorithm choices."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pymcl_b200
from pymcl_b200 import _lib, synth
ctx = _lib.Context(0)
for nq, nm in ((150, 220), (400, 220), (2048, 2048)):
    md, fr = synth.sift_map_and_frames(1, 25, nm, 1, nq)
    m = _lib.MapIndex(ctx, "SIFT", md)
    res = {}
    for name, algo in (("auto", _lib.ALGO_AUTO), ("tc4", _lib.ALGO_TENSOR), ("simt", _lib.ALGO_SIMT)):
        for _ in range(5):
            c = m.match_counts(fr, algo=algo)
        t0 = time.perf_counter()
        for _ in range(50):
            c = m.match_counts(fr, algo=algo)
        res[name] = (time.perf_counter() - t0) / 50 * 1e6
    print("Nq=%d Nm=%d x25 images: " % (nq, nm) + "  ".join("%s %.0f us" % kv for kv in res.items()))
    m.close()
