"""sift_tc4 on small map images (config 2's regime): kernel time vs rows per image."""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import pymcl_b200
from pymcl_b200 import _lib, synth
ctx = _lib.Context(0)
for nm, nq in ((256, 256), (224, 256), (220, 150), (192, 192), (150, 150), (500, 500)):
    map_des, frames = synth.sift_map_and_frames(2, 400, nm, 200, nq)
    m = _lib.MapIndex(ctx, "SIFT", map_des)
    for _ in range(3):
        m.match_counts(frames, algo=_lib.ALGO_TENSOR)
    for k in _lib.PROFILE_SLOTS:
        ctx.profile_read(k)
    ctx.profile_enable(True)
    for _ in range(5):
        m.match_counts(frames, algo=_lib.ALGO_TENSOR)
    ms, n = ctx.profile_read("sift_tc")
    fx, _ = ctx.profile_read("sift_fixup")
    ctx.profile_enable(False)
    ops = 2.0 * nm * nq * 128 * 80000
    print(json.dumps({"rows_per_image": nm, "rows_per_frame": nq, "kernel_ms": round(ms / 5, 4), "fixup_ms": round(fx / 5, 4),
                      "useful_TOPs": round(ops / (ms / 5 * 1e-3) / 1e12, 1)}))
    m.close()
