"""Batched bag-of-words (createRawP's shape): F frames x L locations in one mcl_bow_score call; per-kernel times."""
import json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "oracle"))
import numpy as np
import pymcl_b200
from pymcl_b200 import _lib, synth

F = int(sys.argv[1]) if len(sys.argv) > 1 else 50
L = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ctx = _lib.default_context()
locs = []
for l in range(L):
    rng = np.random.default_rng(500 + l)
    voc = np.clip(synth.sift_like(rng, 10000) + rng.normal(0, 0.3, size=(10000, 128)), 0, 255).astype(np.float32)
    im = rng.random((100, 10000), dtype=np.float32) * (rng.random((100, 10000)) < 0.03)
    im /= np.maximum(np.linalg.norm(im, axis=1, keepdims=True), 1e-9)
    idf = rng.random(10000, dtype=np.float32) + 0.5
    locs.append((im.astype(np.float32), idf, voc))
bow = _lib.BowIndex(ctx, locs, 10000)
rng = np.random.default_rng(9)
frames = [np.clip(np.rint(locs[f % L][2][rng.choice(10000, 300)] + rng.normal(0, 6.0, size=(300, 128))), 0, 255).astype(np.float32)
          for f in range(F)]
slots = ("bow_vq_tc", "bow_resolve", "bow_vq", "bow_score", "ingest")
for mode in (0, 1):
    ctx.set_option(2, mode)
    for _ in range(2):
        s0, _ = bow.score(frames)
    t0 = time.perf_counter()
    for _ in range(3):
        bow.score(frames)
    dt = (time.perf_counter() - t0) / 3
    for k in slots:
        ctx.profile_read(k)
    ctx.profile_enable(True)
    for _ in range(3):
        bow.score(frames)
    r = {k: round(ctx.profile_read(k)[0] / 3, 4) for k in slots}
    ctx.profile_enable(False)
    if mode == 0:
        ref = s0
    print(json.dumps({"frames": F, "locations": L, "simt_only": bool(mode), "e2e_ms": round(dt * 1e3, 3), **r,
                      "frame_locations_per_s": round(F * L / dt), "max_abs_diff_vs_tc": float(np.abs(s0 - ref).max())}))
ctx.set_option(2, 0)
