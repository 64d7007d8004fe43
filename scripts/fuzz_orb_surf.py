"""Fuzz of the ORB (knn + crossCheck) and SURF kernels against the oracle."""
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
    def orb(k):
        s = int(rng.integers(0, 3))
        if s == 0: return synth.orb_like(rng, k)
        if s == 1: return (rng.integers(0, 2, size=(k, 32)) * 255).astype(np.uint8)
        return np.repeat(synth.orb_like(rng, max(1, k // 4)), 4, axis=0)[:k]
    md = [orb(int(rng.integers(0, 600))) for _ in range(int(rng.integers(1, 7)))]
    md = [m if len(m) else None for m in md]
    fr = [orb(int(rng.integers(1, 600))) for _ in range(int(rng.integers(1, 4)))]
    mi = _lib.MapIndex(ctx, "ORB", md)
    k = mi.match_counts(fr, mode=_lib.KNN2_RATIO); c = mi.match_counts(fr, mode=_lib.CROSSCHECK)
    mi.close()
    wk = np.array([[oracle.orb_pair_count_knn(q, m) for m in md] for q in fr]); wc = np.array([[oracle.orb_pair_count_crosscheck(q, m) for m in md] for q in fr])
    if not (np.array_equal(k, wk) and np.array_equal(c, wc)):
        bad += 1; print("ORB FAIL seed", seed)
    sd = [synth.surf_like(rng, int(rng.integers(2, 500))) for _ in range(int(rng.integers(1, 5)))]
    sq = [synth.surf_like(rng, int(rng.integers(1, 400)))]
    n_c = min(len(sq[0]), len(sd[0])) // 2
    if n_c: sq[0][:n_c] = synth.surf_noisy_copy(rng, sd[0][:n_c], float(rng.choice([0.005, 0.02, 0.05])))
    mi = _lib.MapIndex(ctx, "SURF", sd)
    g = mi.match_counts(sq)
    mi.close()
    for i, d in enumerate(sd):
        w, amb = oracle.surf_pair_count(sq[0], d, tie_rel=1e-6)
        if abs(int(g[0, i]) - w) > amb:
            bad += 1; print("SURF FAIL seed", seed, i, g[0, i], w, amb)
print("orb/surf fuzz done:", n, "seeds,", bad, "failures")
