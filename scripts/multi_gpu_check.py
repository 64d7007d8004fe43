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
root = os.path.join(tempfile.gettempdir(), "mcl_multi_gpu_ds")
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
dist.destroy_process_group()
