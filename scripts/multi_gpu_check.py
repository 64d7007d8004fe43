"""Multi-GPU check of the reference-facing drivers (run under torchrun, one rank per GPU):
analyzer('SIFT').createRawP() + processRaw() with the map images sharded over the ranks, frames extracted round-robin,
counts all-gathered over NCCL; rank 0 compares the files with the golden ones written by the unmodified reference."""
import glob, io, contextlib, os, sys, tempfile
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch, torch.distributed as dist
import pymcl_b200
from pymcl_b200 import synth

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
root = os.path.join(os.environ.get("MCL_MULTI_GPU_ROOT") or tempfile.gettempdir(), "mcl_multi_gpu_ds")
if rank == 0:
    synth.make_dataset(root, seed=0, n_locations=7, n_images=25, n_frames=6, view=(160, 120))
dist.barrier()
os.chdir(root)
orig = glob.glob
glob.glob = lambda *a, **k: sorted(orig(*a, **k))
from pymcl_b200.analyze import analyzer
with contextlib.redirect_stdout(io.StringIO()):
    a = analyzer("SIFT", 160, 120)
    a.createRawP()
    dist.barrier()
    a.processRaw()
dist.barrier()
assert a._map.world == world and a._map.hi - a._map.lo < 175
gold = os.path.join(REPO, "tests", "golden")
if rank == 0:
    ok = all(open(f).read() == open(os.path.join(gold, "sift_" + f)).read() for f in ("rawP.txt", "out.txt", "bestGuess.txt"))
    print("MULTI_GPU_OK" if ok else "MULTI_GPU_MISMATCH", "world", world, "shards", a._map.bounds)
dist.barrier()

# colour method: frames' histograms extracted round-robin across the ranks, one chi-squared call per rank
with contextlib.redirect_stdout(io.StringIO()):
    a = analyzer("Color", 160, 120)
    a.createRawP()
    dist.barrier()
    a.processRaw()
dist.barrier()
if rank == 0:
    ok = all(open(f).read() == open(os.path.join(gold, "color_" + f)).read() for f in ("rawP.txt", "out.txt", "bestGuess.txt"))
    print("MULTI_GPU_COLOR_OK" if ok else "MULTI_GPU_COLOR_MISMATCH")
dist.barrier()

# bag of words: index build dealt to the ranks (writeIndices), LOCATIONS sharded for scoring, float32 scores all-gathered;
# rank 0 then scores every location on its own GPU from the same .pkl files and compares the text
import numpy as np, joblib
from pymcl_b200 import _lib
from pymcl_b200.Matcher import Matcher
np.random.seed(5)
with contextlib.redirect_stdout(io.StringIO()):
    m = Matcher("BOW", width=160, height=120)
    m.numWords = 200
    m.writeIndices()
    a = analyzer("BOW", 160, 120)
    a.createRawP()
frames = sorted(orig("cam1_img/*.png"))
qdes = a._extract_many(frames, "SIFT")       # collective: every rank takes part
dist.barrier()
if rank == 0:
    locs = []
    for i in range(7):
        im_features, image_paths, idf, numWords, voc = joblib.load("map/%d.pkl" % i)
        locs.append((im_features, idf, voc))
    bow = _lib.BowIndex(_lib.default_context(), locs, 200)
    scores, _ = bow.score(qdes)
    want = []
    for f in range(len(frames)):
        for i in range(7):
            s = scores[f:f + 1, bow.img_off[i]:bow.img_off[i + 1]]
            want.append([10 * np.max(s), s[0].tolist()])
    a.writeProb(want, "rawP_single.txt", "w")
    ok = open("rawP.txt").read() == open("rawP_single.txt").read() and len(want) == 42
    print("MULTI_GPU_BOW_OK" if ok else "MULTI_GPU_BOW_MISMATCH")
dist.barrier()

# descriptor level, the sharded paths of engine.ShardedMap over NCCL against one GPU holding the whole map:
#  (a) a big batch: every frame uploaded by ONE rank, rows all-gathered over NVLink, counts all-gathered on the device;
#  (b) a small batch with a DOR mask: uploaded whole by every rank;  (c) resident queries (match_counts_device);  (d) ORB.
from pymcl_b200.engine import ShardedMap
ctx = _lib.default_context()
rng = np.random.default_rng(3)
map_des = [synth.sift_like(rng, int(n)) for n in rng.integers(300, 900, size=23)] + [None]
frames = [synth.sift_like(rng, int(n)) for n in rng.integers(700, 1300, size=44)]
for f in range(0, 44, 3):
    src = map_des[(5 * f) % 23]
    frames[f][:200] = synth.sift_noisy_copy(rng, src[:200], 4.0)
assert sum(f.nbytes for f in frames) >= ShardedMap.SHARED_UPLOAD_MIN_BYTES
sm = ShardedMap("SIFT", map_des, ctx=ctx)
got_big = sm.match_counts(frames)
mask = (rng.random((3, 24)) < 0.5).astype(np.uint8)
got_small = sm.match_counts(frames[:3], mask=mask, fill_value=1)
q = _lib.QueryBatch(ctx, "SIFT", frames[:9])
got_dev = sm.match_counts_device(q).cpu().numpy()
q.close()
bad = sm.match_counts  # non-integer rows must be rejected on the sharded path too (checked on the device, reported after the step)
try:
    broken = [f.copy() for f in frames]
    broken[17][3, 3] = 0.5
    sm.match_counts(broken)
    rejected = False
except _lib.MclError:
    rejected = True
omap = [synth.orb_like(rng, int(n)) for n in rng.integers(100, 600, size=13)]
oq = [synth.orb_like(rng, 500) for _ in range(5)]
n_pl = min(150, len(omap[4]))
oq[1][:n_pl] = synth.orb_flip(rng, omap[4][:n_pl], 25)
so = ShardedMap("ORB", omap, ctx=ctx)
got_orb = [so.match_counts(oq, mode=m) for m in (_lib.KNN2_RATIO, _lib.CROSSCHECK)]
if rank == 0:
    whole = _lib.MapIndex(ctx, "SIFT", map_des)
    want_big = whole.match_counts(frames)
    checks = {"big": np.array_equal(got_big, want_big), "planted": bool(got_big[0, 0] >= 150),
              "small_masked": np.array_equal(got_small, np.where(mask > 0, want_big[:3], 1)),
              "resident": np.array_equal(got_dev, want_big[:9]), "rejected": rejected}
    ok = all(checks.values())
    if not ok:
        print("sharded checks:", checks, "differing cells (big):", int((got_big != want_big).sum()), "of", got_big.size,
              "first rows got", got_big[:2].tolist(), "want", want_big[:2].tolist())
    whole.close()
    print("MULTI_GPU_SHARDED_UPLOAD_OK" if ok else "MULTI_GPU_SHARDED_UPLOAD_MISMATCH", "shards", sm.bounds)
    wo = _lib.MapIndex(ctx, "ORB", omap)
    want_orb = [wo.match_counts(oq, mode=m) for m in (_lib.KNN2_RATIO, _lib.CROSSCHECK)]
    ok = all(np.array_equal(g, w) for g, w in zip(got_orb, want_orb))
    if not ok:
        print("orb:", [(int((g != w).sum()), g.shape, w.shape) for g, w in zip(got_orb, want_orb)], "shards", so.bounds,
              "got", got_orb[1][:2].tolist(), "want", want_orb[1][:2].tolist())
    wo.close()
    print("MULTI_GPU_ORB_OK" if ok else "MULTI_GPU_ORB_MISMATCH")
sm.close()
so.close()
dist.barrier()
dist.destroy_process_group()
