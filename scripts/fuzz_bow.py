"""Fuzz of the BOW word assignment: tensor-core path vs float32 SIMT path vs the float64 oracle."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "oracle"))
import numpy as np
import pymcl_b200
from pymcl_b200 import _lib, synth
import oracle
ctx = _lib.default_context()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
bad = 0
for seed in range(n):
    rng = np.random.default_rng(seed)
    L = int(rng.integers(1, 5)); W = int(rng.choice([64, 200, 513, 1500]))
    locs = []
    for l in range(L):
        rows = int(rng.integers(1, W + 1))
        style = int(rng.integers(0, 3))
        if style == 0:
            voc = np.clip(synth.sift_like(rng, rows) + rng.normal(0, 0.4, size=(rows, 128)), 0, 255)
        elif style == 1:
            voc = rng.uniform(0, 255.9, size=(rows, 128))
        else:
            voc = np.clip(synth.sift_like(rng, max(2, rows // 8))[rng.integers(0, max(2, rows // 8), size=rows)] + rng.normal(0, 0.01, size=(rows, 128)), 0, 255)
        locs.append((np.zeros((1, W), np.float32), np.ones(W, np.float32), voc.astype(np.float32)))
    nq = int(rng.integers(1, 700))
    q = synth.sift_like(rng, nq) if rng.random() < 0.5 else np.rint(locs[0][2][rng.integers(0, len(locs[0][2]), size=nq)] + rng.normal(0, 3, size=(nq, 128))).clip(0, 255).astype(np.float32)
    bow = _lib.BowIndex(ctx, locs, W)
    _, w_tc = bow.score([q], want_words=True)
    ctx.set_option(2, 1)
    _, w_simt = bow.score([q], want_words=True)
    ctx.set_option(2, 0)
    bow.close()
    for l in range(L):
        ow, gap = oracle.bow_words(q, locs[l][2])
        for name, got in (("tc", w_tc[l]), ("simt", w_simt[l])):
            d = got != ow
            if not np.all(gap[d] < 1e-5):
                bad += 1
                print("FAIL seed", seed, "loc", l, name, "rows", np.nonzero(d & (gap >= 1e-5))[0][:5], "gap", gap[d & (gap >= 1e-5)][:5])
print("bow fuzz done:", n, "seeds,", bad, "failures")
