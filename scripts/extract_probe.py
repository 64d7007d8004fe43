"""Frame extraction (cv2 setQuery + SIFT) on the thread pool: cv2's own threading vs one cv2 thread per worker."""
import glob, os, sys, tempfile, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import cv2
import pymcl_b200
from pymcl_b200 import synth
root = tempfile.mkdtemp(prefix="mcl_ext_")
synth.make_dataset(root, seed=2, n_locations=1, n_images=25, n_frames=64, view=(320, 240), n_shapes=400)
os.chdir(root)
open("commands.txt", "a").close()
from pymcl_b200.analyze import analyzer
a = analyzer("SIFT", 320, 240, numLocations=1)
frames = sorted(glob.glob("cam1_img/*.png"))
print("cores", os.cpu_count(), "cv2 threads", cv2.getNumThreads(), "frame", cv2.imread(frames[0]).shape)
for nthreads in (cv2.getNumThreads(), 1, 2):
    cv2.setNumThreads(nthreads)
    for workers in (16, 8, 32):
        os.environ["MCL_EXTRACT_THREADS"] = str(workers)
        t0 = time.perf_counter(); d = a._extract_many(frames); dt = time.perf_counter() - t0
        print("cv2 threads %2d  pool workers %2d: %.3f s for %d frames (%.1f ms/frame)" % (nthreads, workers, dt, len(frames), dt / len(frames) * 1e3))
