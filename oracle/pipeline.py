"""CPU restatement of the reference's sequence drivers (analyze.py:64-232) on top of oracle.py.

TEST INFRASTRUCTURE ONLY (see oracle.py).  Given a dataset directory laid out the way the reference expects
(map/<l>/angleNNN.png, cam1_img/NNNN.png, commands.txt) it produces the same rawP / out / bestGuess lists as the
unmodified reference run through oracle/ref_shims.py with exact search -- pinned byte-for-byte against
tests/golden/*.txt by tests/test_oracle.py.  Frames are taken in sorted order (the golden run pinned glob order).
"""
from __future__ import annotations

import glob
import os

import cv2
import numpy as np

import oracle

EXT = ".png"


def angle_paths(root, loc, n_images=25, step=15):
    return [os.path.join(root, "map", str(loc), "angle" + str(i * step).zfill(3) + EXT) for i in range(n_images)]


def extractor(kind):
    return cv2.SIFT_create() if kind == "SIFT" else cv2.ORB_create()


def map_descriptors(root, kind, n_locations=7, n_images=25, step=15):
    """Matcher.createFeatureIndex (Matcher.py:147-163): raw imread, no preprocessing."""
    ex = extractor(kind)
    out = []
    for loc in range(n_locations):
        out.append([ex.detectAndCompute(cv2.imread(p), None)[1] for p in angle_paths(root, loc, n_images, step)])
    return out


def frame_paths(root):
    return sorted(glob.glob(os.path.join(root, "cam1_img", "*" + EXT)))


def query_descriptors(path, kind, w, h):
    """Matcher.setQuery (Matcher.py:48-51) + the extractor call of *Match."""
    img = cv2.imread(path)
    img = cv2.bilateralFilter(img, 9, 75, 75)
    img = cv2.resize(img, (w, h))
    return extractor(kind).detectAndCompute(img, None)[1]


def pair_count(kind, orb_mode, q, m):
    if kind == "SIFT":
        return oracle.sift_pair_count(q, m)
    if orb_mode == "crosscheck":
        return oracle.orb_pair_count_crosscheck(q, m)
    return oracle.orb_pair_count_knn(q, m)


def read_commands(root):
    d = {}
    for line in open(os.path.join(root, "commands.txt")).read().split("\n")[:-1]:
        d[line[:4]] = str(line[-1])
    return d


def create_raw_p(root, kind, w, h, orb_mode="knn", n_locations=7, n_images=25, step=15):
    """analyzer.createRawP (analyze.py:64-96) -> list of [total, [p...]] (frames x locations).  n_locations / n_images / step
    generalise the reference's 7 / 25 / 15 (analyze.py:35, Matcher.py:309)."""
    mdes = map_descriptors(root, kind, n_locations, n_images, step)
    p = []
    for fp in frame_paths(root):
        q = query_descriptors(fp, kind, w, h)
        for loc in range(n_locations):
            counts = [pair_count(kind, orb_mode, q, m) for m in mdes[loc]]
            total, probs = oracle.run_from_counts(counts)
            p.append([total, probs])
    return p


def process_raw(root, raw_p_file, n_locations=7, n_images=25, step=15):
    """analyzer.processRaw (analyze.py:98-146) -> (blurP, bestGuess).  Keeps the reference's aliasing: accountCommand
    of frame t+1 mutates the inner lists already appended to blurP for frame t."""
    commands = read_commands(root)
    prob = oracle.read_prob(raw_p_file, n_locations)
    previous = oracle.initial_probs(n_locations, n_images)
    blur_p, best = [], []
    for fp in frame_paths(root):
        stem = os.path.basename(fp).replace(EXT, "")
        blur = oracle.laplacian_var(cv2.imread(fp, 0))
        adjusted, bg = oracle.filter_step(commands[stem], previous, prob[stem], blur, step)
        best.append([bg[0], bg[1]])
        blur_p.extend(adjusted)
        previous = adjusted
    return blur_p, best


def opt_p(root, kind, w, h, orb_mode="knn", n_locations=7, n_images=25, step=15):
    """analyzer.optP (analyze.py:148-232): DOR location window +-2 and angle window +-30 degrees."""
    commands = read_commands(root)
    mdes = map_descriptors(root, kind, n_locations, n_images, step)
    previous = oracle.initial_probs(n_locations, n_images)
    best_angle = best_circle = None
    blur_p, best = [], []
    for fp in frame_paths(root):
        q = query_descriptors(fp, kind, w, h)
        window = oracle.opt_window(best_angle, n_images, step)
        p = []
        for loc in range(n_locations):
            if best_circle is None or (best_circle - 2 <= loc <= best_circle + 2):
                counts = [pair_count(kind, orb_mode, q, m) if wd else 1 for m, wd in zip(mdes[loc], window)]
                total, probs = oracle.run_from_counts(counts)
                p.append([total, probs])
            else:
                p.append([1, [1 / (3 * n_images)] * n_images])
        stem = os.path.basename(fp).replace(EXT, "")
        blur = oracle.laplacian_var(cv2.imread(fp, 0))
        adjusted, bg = oracle.filter_step(commands[stem], previous, p, blur, step)
        best_circle, best_angle = bg
        best.append([bg[0], bg[1]])
        blur_p.extend(adjusted)
        previous = adjusted
    return blur_p, best


# ---- colour-histogram method (Matcher.py:67-99, search.py) -------------------------------------------------------
def color_histogram(image, bins=(8, 8, 8)):
    """Matcher.createHistogram (Matcher.py:67-74)."""
    hist = cv2.calcHist([image], [0, 1, 2], None, list(bins), [0, 256, 0, 256, 0, 256])
    hist = cv2.normalize(hist, hist, 0, 255, cv2.NORM_MINMAX)
    return hist.flatten()


def color_index(root, loc):
    """Matcher.createColorIndex (Matcher.py:76-88): {filename: histogram} of one location."""
    idx = {}
    for pth in sorted(glob.glob(os.path.join(root, "map", str(loc), "*" + EXT))):
        idx[os.path.basename(pth)] = color_histogram(cv2.imread(pth))
    return idx


def color_query(path, w, h):
    """Matcher.setQuery (Matcher.py:48-51) followed by createHistogram of the filtered, resized frame."""
    img = cv2.imread(path)
    img = cv2.bilateralFilter(img, 9, 75, 75)
    img = cv2.resize(img, (w, h))
    return color_histogram(img)


def color_location(qh, index, data_dir):
    """colorSearch (Matcher.py:89-99 + search.py:7-20) and the colour branch of run()."""
    results = sorted([(oracle.chi2_literal(feat, qh), name) for name, feat in index.items()])
    return oracle.color_run(results, data_dir)


def create_raw_p_color(root, w, h, n_locations=7):
    """analyzer('Color').createRawP (analyze.py:64-96)."""
    idx = [color_index(root, loc) for loc in range(n_locations)]
    p = []
    for fp in frame_paths(root):
        qh = color_query(fp, w, h)
        for loc in range(n_locations):
            total, probs = color_location(qh, idx[loc], "map/" + str(loc))
            p.append([total, probs])
    return p


def opt_p_color(root, w, h, n_locations=7):
    """analyzer('Color').optP (analyze.py:148-232).  The colour branch of optRun ignores the angle window; run()'s
    np.float32 values flow into the filter, whose Python operators then follow numpy's scalar promotion (float32 with
    Python floats as weak scalars) exactly as in the reference."""
    commands = read_commands(root)
    idx = [color_index(root, loc) for loc in range(n_locations)]
    previous = oracle.initial_probs(n_locations)
    best_circle = None
    blur_p, best = [], []
    for fp in frame_paths(root):
        qh = color_query(fp, w, h)
        p = []
        for loc in range(n_locations):
            if best_circle is None or (best_circle - 2 <= loc <= best_circle + 2):
                total, probs = color_location(qh, idx[loc], "map/" + str(loc))
                p.append([total, probs])
            else:
                p.append([1, [1 / 75] * 25])
        stem = os.path.basename(fp).replace(EXT, "")
        # ndarray.var() returns np.float64 (a strong type): with blur <= 200 the blur weight promotes the float32 values
        # to float64, with blur > 200 the weight is the Python float 0.85 and they stay float32
        blur = np.float64(oracle.laplacian_var(cv2.imread(fp, 0)))
        adjusted, bg = oracle.filter_step(commands[stem], previous, p, blur)
        best_circle = bg[0]
        best.append([bg[0], bg[1]])
        blur_p.extend(adjusted)
        previous = adjusted
    return blur_p, best
