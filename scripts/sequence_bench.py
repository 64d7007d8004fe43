"""Sequence-level benchmark through the reference-facing API (BASELINE config 2's shape at the reference's hard-coded
7 locations x 25 images): analyzer('SIFT', w, h).createRawP() + processRaw() on a synthetic sequence, with the oracle
pipeline (cv2 brute force on the host) timed on a few frames beside it.
    python scripts/sequence_bench.py [n_frames] [w] [h]"""
import glob, json, os, sys, tempfile, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "oracle"))
import numpy as np
import pymcl_b200
from pymcl_b200 import synth

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 200
w = int(sys.argv[2]) if len(sys.argv) > 2 else 320
h = int(sys.argv[3]) if len(sys.argv) > 3 else 240
root = tempfile.mkdtemp(prefix="mcl_seq_")
synth.make_dataset(root, seed=2, n_locations=7, n_images=25, n_frames=n_frames, view=(w, h), n_shapes=400)
os.chdir(root)
orig = glob.glob
glob.glob = lambda *a, **k: sorted(orig(*a, **k))
import io, contextlib
from pymcl_b200.analyze import analyzer
res = {"config": "analyzer('SIFT',%d,%d): %d-frame sequence vs 7 locations x 25 images" % (w, h, n_frames)}
with contextlib.redirect_stdout(io.StringIO()):
    a = analyzer("SIFT", w, h)
    t0 = time.perf_counter(); a.createIndex(); t1 = time.perf_counter()
    a.createRawP = a.createRawP  # (createRawP re-runs createIndex like the reference; timed as a whole below)
    t2 = time.perf_counter(); a.createRawP(); t3 = time.perf_counter()
    a.processRaw(); t4 = time.perf_counter()
with contextlib.redirect_stdout(io.StringIO()):
    b = analyzer("SIFT", w, h)
    t5 = time.perf_counter(); b.optP(); t6 = time.perf_counter()
res["optP_s(incl. index)"] = t6 - t5
res.update({"createIndex_s": t1 - t0, "createRawP_s(incl. index)": t3 - t2, "processRaw_s": t4 - t3,
            "image_matches": n_frames * 175, "image_matches_per_s_createRawP": n_frames * 175 / (t3 - t2)})
# oracle pipeline on the host for the first few frames (cv2 extraction + cv2 brute force + numpy filter)
import cv2, oracle, pipeline
k = min(4, n_frames)
t0 = time.perf_counter()
mdes = pipeline.map_descriptors(".", "SIFT")
t1 = time.perf_counter()
cnt = 0
for fp in pipeline.frame_paths(".")[:k]:
    q = pipeline.query_descriptors(fp, "SIFT", w, h)
    for loc in range(7):
        for m in mdes[loc]:
            oracle.cv2_l2_pair_count(q, m); cnt += 1
t2 = time.perf_counter()
res.update({"cpu_map_extract_s": t1 - t0, "cpu_frames": k, "cpu_image_matches_per_s": cnt / (t2 - t1), "cpu_threads": cv2.getNumThreads()})
print(json.dumps(res))
