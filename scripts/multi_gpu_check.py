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
if rank == 0:
    gold = os.path.join(REPO, "tests", "golden")
    ok = all(open(f).read() == open(os.path.join(gold, "sift_" + f)).read() for f in ("rawP.txt", "out.txt", "bestGuess.txt"))
    print("MULTI_GPU_OK" if ok else "MULTI_GPU_MISMATCH", "world", world, "shards", a._map.bounds)
dist.barrier()
dist.destroy_process_group()
