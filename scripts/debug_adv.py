import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "oracle"))
import numpy as np
import pymcl_b200
from pymcl_b200 import _lib, synth
import oracle
ctx = _lib.Context(0)
rng = np.random.default_rng(41)
def wild(n):
    base = synth.sift_like(rng, n)
    scale = rng.choice([0.05, 0.2, 0.5, 1.0], size=(n, 1))
    return np.clip(np.rint(base * scale), 0, 255).astype(np.float32)
map_des = [wild(n) for n in (700, 90, 333, 64, 1500)]
frames = [wild(600), wild(257)]
for f, (i, n) in enumerate(((0, 300), (4, 200))):
    a = map_des[i][rng.integers(0, len(map_des[i]), size=n)]
    b = map_des[i][rng.integers(0, len(map_des[i]), size=n)]
    w = rng.uniform(0.55, 0.75, size=(n, 1))
    frames[f][:n] = np.clip(np.rint(w * a + (1 - w) * b), 0, 255)
m = _lib.MapIndex(ctx, "SIFT", map_des)
for name, algo in (("simt", 2), ("tc4", 1), ("tc3", 5), ("tc2", 4), ("tc1", 3)):
    print(name, m.match_counts(frames, algo=algo).tolist())
# row-level: frame 1 vs image 4
q = frames[1]; img = map_des[4]
d2 = oracle.l2sq_int(q, img)
d0, d1 = oracle.top2_smallest(d2)
want = oracle.ratio_pass_from_d2(d0, d1)
rows = [q[i:i + 1] for i in range(len(q))]
got = m.match_counts(rows, algo=1)[:, 4]
bad = np.nonzero(got != want.astype(int))[0]
print("bad rows", bad, "want", want[bad], "got", got[bad])
nrm = (img.astype(np.int64) ** 2).sum(1)
order = np.argsort(nrm, kind="stable")
for r in bad[:3]:
    keys = -(d2[r] - (q[r].astype(np.int64) ** 2).sum())   # 2 q.m - |m|^2
    ks = keys[order]                                        # sorted-by-norm layout
    j1, j2 = np.argsort(-ks)[:2]
    print("row", r, "Q", int((q[r].astype(np.int64) ** 2).sum()), "best", int(ks[j1]), "at sorted pos", int(j1), "tile", int(j1) // 64,
          "second", int(ks[j2]), "pos", int(j2), "tile", int(j2) // 64, "d0", int(d0[r]), "d1", int(d1[r]))
    A = (ks + nrm[order]) // 2
    for t in range((len(ks) + 63) // 64):
        sl = slice(t * 64, min(len(ks), (t + 1) * 64))
        gm = int(A[sl].max()); nmin = int(nrm[order][sl].min()); nmax = int(nrm[order][sl].max())
        print("   tile", t, "hi", 2 * gm - nmin, "lo", 2 * gm - nmax, "true best", int(ks[sl].max()))
