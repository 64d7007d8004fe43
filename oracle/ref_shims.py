"""Import the UNMODIFIED reference modules (Matcher.py / analyze.py under /root/reference) in a
modern environment.  Build-container only: /root/reference does not exist on the GPU box, so
nothing at test/bench run time may call this — it is used by oracle/make_golden.py to produce
the committed fixtures under tests/golden/ and by tests that skip when the checkout is absent.

Shims:
  1. matplotlib is absent           -> empty stub modules (only used with display_results=True)
  2. sklearn.externals.joblib       -> the standalone joblib package
  3. cv2.xfeatures2d.SIFT_create    -> cv2.SIFT_create (SURF cannot be shimmed: non-free is OFF)
  4. numpy >= 2 float repr          -> legacy print options so str(list) matches the 2016 format
  5. exact search (north_star, D1)  -> cv2.FlannBasedMatcher replaced by BFMatcher(NORM_L2)
  6. ORB (D2): the source raises NameError; `patch_orb(mode)` restates its 4 lines with either
     knnMatch(k=2)+ratio lambda (north_star) or the bytecode's len(bf.match(...)).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MCL_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "Matcher.py"))


def load(exact_search=True):
    """Returns (Matcher_module, analyze_module)."""
    import cv2
    import numpy as np

    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    import sklearn  # noqa: F401
    import joblib
    if "sklearn.externals" not in sys.modules:
        ext = types.ModuleType("sklearn.externals")
        sys.modules["sklearn.externals"] = ext
        sklearn.externals = ext
    sys.modules["sklearn.externals"].joblib = joblib
    sys.modules["sklearn.externals.joblib"] = joblib
    if not hasattr(cv2, "xfeatures2d") or not hasattr(cv2.xfeatures2d, "SIFT_create"):
        cv2.xfeatures2d = types.SimpleNamespace(SIFT_create=cv2.SIFT_create)
    np.set_printoptions(legacy="1.25")
    if exact_search and not getattr(cv2, "_mcl_flann_patched", False):
        cv2._mcl_orig_flann = cv2.FlannBasedMatcher
        cv2.FlannBasedMatcher = lambda *a, **k: cv2.BFMatcher(cv2.NORM_L2)
        cv2._mcl_flann_patched = True
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    for name in ("search", "Matcher", "analyze"):
        mod = sys.modules.get(name)
        if mod is not None and not str(getattr(mod, "__file__", "")).startswith(REFERENCE_ROOT):
            del sys.modules[name]
    ref_matcher = importlib.import_module("Matcher")
    ref_analyze = importlib.import_module("analyze")
    return ref_matcher, ref_analyze


def unload():
    """Drop the reference modules from sys.modules / sys.path so the product's same-named drop-in
    modules can be imported afterwards."""
    for name in ("search", "Matcher", "analyze"):
        sys.modules.pop(name, None)
    while REFERENCE_ROOT in sys.path:
        sys.path.remove(REFERENCE_ROOT)


def patch_orb(ref_matcher, mode="knn"):
    """Restates Matcher.py:169-194 without its NameError."""
    import cv2

    def ORBMatch(self, imagePath, display_results=False):
        orb = cv2.ORB_create()
        kp1, des1 = orb.detectAndCompute(self.image, None)
        kp2, des2 = self.index[imagePath]
        if mode == "crosscheck":
            bf = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True)
            return len(bf.match(des1, des2))
        bf = cv2.BFMatcher(cv2.NORM_HAMMING)
        matches = bf.knnMatch(des1, des2, k=2)
        filtered = list(filter(lambda x: x[0].distance < 0.7 * x[1].distance, matches))
        return len(filtered)

    ref_matcher.Matcher.ORBMatch = ORBMatch
