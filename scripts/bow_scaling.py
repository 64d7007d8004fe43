"""Weak-scaling run of the bag-of-words path (BASELINE config 5: 256 locations x 100 images at 8 GPUs = 32 locations per GPU):
ShardedBow scores a batch of frames against every location -- locations sharded over the ranks, float32 scores all-gathered.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bow_scaling.py [frames] [locations per GPU]"""
import json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch, torch.distributed as dist
import pymcl_b200
from pymcl_b200 import _lib, engine, synth

F = int(sys.argv[1]) if len(sys.argv) > 1 else 50
LPG = int(sys.argv[2]) if len(sys.argv) > 2 else 32
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank = dist.get_rank() if world > 1 else 0
L = LPG * world


def loader(l):
    rng = np.random.default_rng(500 + l)
    voc = np.clip(synth.sift_like(rng, 10000) + rng.normal(0, 0.3, size=(10000, 128)), 0, 255).astype(np.float32)
    im = rng.random((100, 10000), dtype=np.float32) * (rng.random((100, 10000)) < 0.03)
    im /= np.maximum(np.linalg.norm(im, axis=1, keepdims=True), 1e-9)
    return im.astype(np.float32), rng.random(10000, dtype=np.float32) + 0.5, voc


ctx = _lib.default_context()
bow = engine.ShardedBow(L, loader, num_words=10000, ctx=ctx)
rng = np.random.default_rng(9)
base = loader(0)[2]
frames = [np.clip(np.rint(base[rng.choice(10000, 300)] + rng.normal(0, 6.0, size=(300, 128))), 0, 255).astype(np.float32) for _ in range(F)]
for _ in range(2):
    s = bow.score(frames)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
reps = 5
t0 = time.perf_counter()
for _ in range(reps):
    s = bow.score(frames)
torch.cuda.synchronize()
dt = torch.tensor([(time.perf_counter() - t0) / reps], device="cuda")
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
assert s.shape == (F, L * 100)
if rank == 0:
    print(json.dumps({"config": "cfg5 BOW weak scaling: %d frames x 300 rows vs %d locations x 100 images (W=10000), %d per GPU" % (F, L, LPG),
                      "n_gpus": world, "ms_per_batch_max_over_ranks": float(dt[0]) * 1e3,
                      "frame_locations_per_s": F * L / float(dt[0]), "scores_per_s": F * L * 100 / float(dt[0])}))
if world > 1:
    dist.destroy_process_group()
