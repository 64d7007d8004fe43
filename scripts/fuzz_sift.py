"""Long fuzz run of the SIFT kernels (development aid): python scripts/fuzz_sift.py [n_seeds]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "oracle")); sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np
import pymcl_b200
from pymcl_b200 import _lib
import test_gpu_parity as T
ctx = _lib.default_context()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
bad = 0
for seed in range(100, 100 + n):
    try:
        T.test_sift_fuzz_all_generations_equal_exact_kernel(ctx, seed)
    except AssertionError as e:
        bad += 1
        print("FAIL seed", seed, str(e)[:200])
print("fuzz done:", n, "seeds,", bad, "failures, rescans", ctx.info()["sift_rescans"])
