"""CPU oracle for the descriptor-matching hot path of zhangxingshuo/py-mcl.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``py-mcl_b200/``) may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs do.

It is a numpy restatement of what the reference computes on this path.  All heavy arithmetic of
the reference lives in third-party wheels that are not vendored under /root/reference (there is no
requirements file, so no pinned versions): OpenCV (``cv2.BFMatcher`` / ``FlannBasedMatcher`` /
``Laplacian``; 4.13.0 in this image), SciPy (``scipy.cluster.vq.vq``; 1.18.1), scikit-learn
(``preprocessing.normalize``; 1.9.0) and NumPy (2.3.5).  Their published algorithms are restated
here in plain numpy; ``oracle/make_golden.py`` + ``tests/test_oracle.py`` pin every restatement
against (a) the third-party routine itself on seeded inputs, (b) the unmodified reference modules
run through import shims in the build container (``oracle/ref_shims.py``; outputs committed under
``tests/golden/``) and (c) the known-answer vectors in SURVEY.md §8(c).

Reference call sites followed (file:line relative to the reference checkout):
  * ``SIFTMatch``   Matcher.py:262-291  knnMatch(k=2) + ratio lambda Matcher.py:280
  * ``SURFMatch``   Matcher.py:196-230  same on 64-d float descriptors (ratio lambda :219)
  * ``ORBMatch``    Matcher.py:169-194  Hamming brute force (crossCheck, :180-182)
  * ``run``/``optRun``  Matcher.py:305-338 / 340-388
  * ``BOWMatch``    Matcher.py:232-259  vq + histogram + idf + L2 normalise + dot
  * ``Laplacian``   analyze.py:441-445
  * ``accountCommand``/``prevWeight``/``probUpdate``  analyze.py:312-336 / 278-310 / 239-276
  * ``writeProb``/``readProb``  analyze.py:351-356 / 381-398
"""
from __future__ import annotations

import math

import numpy as np

RATIO = 0.7  # Matcher.py:219,280


# --------------------------------------------------------------------------------------------
# L2 (SIFT / SURF) knn-2 + ratio test
# --------------------------------------------------------------------------------------------
def _as_int_desc(des):
    """SIFT descriptors are float32 but integer valued in [0,255]."""
    a = np.asarray(des)
    ai = a.astype(np.int64)
    if not np.array_equal(ai, a):
        raise ValueError("descriptor matrix is not integer valued")
    return ai


def l2sq_int(des1, des2):
    """Exact integer squared distances, (Nq, Nm) int64."""
    a = _as_int_desc(des1)
    b = _as_int_desc(des2)
    na = (a * a).sum(1)[:, None]
    nb = (b * b).sum(1)[None, :]
    # the dot products through float64 BLAS: entries are integers <= 255, so every product and partial sum (<= 128 * 255^2)
    # is an integer far below 2^53 and the result is exact whatever the summation order
    dots = (a.astype(np.float64) @ b.astype(np.float64).T)
    return na + nb - 2 * np.rint(dots).astype(np.int64)


def ratio_pass_from_d2(d0sq, d1sq, ratio=RATIO):
    """cv2 returns sqrtf(exact d2) as float32; the reference lambda (Matcher.py:280) compares
    ``m.distance < 0.7*n.distance`` as Python floats (float64)."""
    d0 = np.sqrt(np.asarray(d0sq, dtype=np.float32)).astype(np.float64)
    d1 = np.sqrt(np.asarray(d1sq, dtype=np.float32)).astype(np.float64)
    return d0 < ratio * d1


def top2_smallest(d):
    """Two smallest values per row (with multiplicity), d is (Nq, Nm>=2)."""
    part = np.partition(d, 1, axis=1)[:, :2]
    part.sort(axis=1)
    return part[:, 0], part[:, 1]


def sift_pair_count(des1, des2, ratio=RATIO):
    """len(good) of SIFTMatch with exact (brute force) 2-NN.  Nm<2 or empty query -> 0
    (the reference raises IndexError / cv2 assertion there; SURVEY §5 defines count 0)."""
    if des1 is None or des2 is None or len(des1) == 0 or len(des2) < 2:
        return 0
    d2 = l2sq_int(des1, des2)
    d0, d1 = top2_smallest(d2)
    return int(ratio_pass_from_d2(d0, d1, ratio).sum())


def l2sq_f32(des1, des2, block=64):
    """Float descriptors (SURF): direct sum of squared differences, accumulated in float64 and
    reported in float64.  cv2 accumulates the same sum in float32 SIMD lanes (1.1e-7 rel)."""
    a = np.asarray(des1, dtype=np.float64)
    b = np.asarray(des2, dtype=np.float64)
    out = np.empty((a.shape[0], b.shape[0]), dtype=np.float64)
    for s in range(0, a.shape[0], block):
        diff = a[s:s + block, None, :] - b[None, :, :]
        out[s:s + block] = (diff * diff).sum(2)
    return out


def surf_pair_count(des1, des2, ratio=RATIO, tie_rel=1e-6):
    """Returns (count, n_ambiguous).  A row is ambiguous when |d0 - ratio*d1| <= tie_rel*d1, i.e.
    fp32 rounding of the distance may flip the ratio test (north_star tolerance)."""
    if des1 is None or des2 is None or len(des1) == 0 or len(des2) < 2:
        return 0, 0
    d2 = l2sq_f32(des1, des2)
    d0s, d1s = top2_smallest(d2)
    d0 = np.sqrt(d0s.astype(np.float32)).astype(np.float64)
    d1 = np.sqrt(d1s.astype(np.float32)).astype(np.float64)
    good = d0 < ratio * d1
    amb = np.abs(d0 - ratio * d1) <= tie_rel * np.maximum(d1, 1e-30)
    return int(good.sum()), int(amb.sum())


# --------------------------------------------------------------------------------------------
# Hamming (ORB)
# --------------------------------------------------------------------------------------------
_POP8 = np.array([bin(i).count("1") for i in range(256)], dtype=np.int32)


def hamming_matrix(des1, des2):
    a = np.asarray(des1, dtype=np.uint8)
    b = np.asarray(des2, dtype=np.uint8)
    x = a[:, None, :] ^ b[None, :, :]
    return _POP8[x].sum(2)


def orb_pair_count_knn(des1, des2, ratio=RATIO):
    """north_star ORB mode: BFMatcher(NORM_HAMMING).knnMatch(k=2) + the ratio lambda."""
    if des1 is None or des2 is None or len(des1) == 0 or len(des2) < 2:
        return 0
    d = hamming_matrix(des1, des2)
    d0, d1 = top2_smallest(d)
    return int((d0.astype(np.float64) < ratio * d1.astype(np.float64)).sum())


def orb_pair_count_crosscheck(des1, des2):
    """The only ORB behaviour the authors ever ran (stale bytecode):
    len(BFMatcher(NORM_HAMMING, crossCheck=True).match(des1, des2)) = mutual first-index NNs."""
    if des1 is None or des2 is None or len(des1) == 0 or len(des2) == 0:
        return 0
    d = hamming_matrix(des1, des2)
    row = d.argmin(1)  # first index on ties
    col = d.argmin(0)
    return int((col[row] == np.arange(len(row))).sum())


# --------------------------------------------------------------------------------------------
# run() / optRun()
# --------------------------------------------------------------------------------------------
def run_from_counts(counts):
    """Matcher.py:320-338: total = sum (0 -> 1); p_k = count_k / total (int/int true division)."""
    counts = [int(c) for c in counts]
    total = sum(counts)
    if total == 0:
        total = 1
    return total, [c / total for c in counts]


def opt_window(best_angle_index, n_images=25, step=15):
    """Matcher.py:340-370: list of bools, True where the image is matched, False -> count 1."""
    if best_angle_index is None:
        return [True] * n_images
    best = best_angle_index * step
    lower, upper = best - 30, best + 30
    full = (n_images - 1) * step  # 360 for 25 images
    out = []
    for k in range(n_images):
        i = k * step
        if lower >= 0 and upper <= full:
            out.append(lower <= i <= upper)
        else:
            out.append(i >= lower % full or i <= upper % full)
    return out


# --------------------------------------------------------------------------------------------
# BOW (Matcher.py:232-259)
# --------------------------------------------------------------------------------------------
def bow_words(des, voc):
    """scipy.cluster.vq.vq restated: argmin_j ||o_i - c_j|| (lowest index on ties), evaluated in
    float64 so that it is the exact answer fp32 implementations approximate.  Also returns the
    relative gap between best and second best squared distance (near-tie detector)."""
    o = np.asarray(des, dtype=np.float64)
    c = np.asarray(voc, dtype=np.float64)
    d2 = (o * o).sum(1)[:, None] - 2.0 * (o @ c.T) + (c * c).sum(1)[None, :]
    words = d2.argmin(1).astype(np.int32)
    if c.shape[0] >= 2:
        p = np.partition(d2, 1, axis=1)[:, :2]
        p.sort(axis=1)
        gap = (p[:, 1] - p[:, 0]) / np.maximum(p[:, 1], 1e-30)
    else:
        gap = np.ones(len(o))
    return words, gap


def bow_score_from_words(words, idf, im_features, num_words):
    """Matcher.py:249-258 in float32, as the reference does."""
    test = np.zeros((1, num_words), "float32")
    for w in words:
        test[0][w] += 1
    test = test * idf
    nrm = np.sqrt((test.astype(np.float64) ** 2).sum())  # sklearn normalize(norm='l2')
    if nrm > 0:
        test = (test / np.float32(nrm)).astype(np.float32)
    return np.dot(test, np.asarray(im_features, dtype=np.float32).T)


# --------------------------------------------------------------------------------------------
# colour histograms: chi-squared search (search.py:7-26) and run()'s colour branch (Matcher.py:328-336)
# --------------------------------------------------------------------------------------------
def chi2_literal(hist_a, hist_b, eps=1e-10):
    """search.py:22-26 as written: a Python list of per-bin terms summed by np.sum.  With float32 histograms every
    term is an np.float32 (the Python float eps is a weak scalar under NEP 50) and np.sum reduces a float32 array."""
    return 0.5 * np.sum([((a - b) ** 2) / (a + b + eps) for (a, b) in zip(hist_a, hist_b)])


def chi2_matrix(q_hists, m_hists, eps=1e-10):
    """The same arithmetic for all (query, map) pairs, vectorised per pair: elementwise float32 ops give the same terms
    and np.sum over the contiguous float32 row uses the same pairwise summation as np.sum of the list."""
    q = np.asarray(q_hists, dtype=np.float32)
    m = np.asarray(m_hists, dtype=np.float32)
    e = np.float32(eps)
    out = np.zeros((q.shape[0], m.shape[0]), dtype=np.float32)
    for i in range(q.shape[0]):
        for j in range(m.shape[0]):
            a, b = m[j], q[i]
            out[i, j] = np.float32(0.5) * np.sum(((a - b) ** 2) / (a + b + e))
    return out


def pairwise_sum_f32(x):
    """numpy's float32 pairwise summation spelled out (blocks <= 128 with 8 interleaved accumulators, split points rounded
    down to multiples of 8): the order of additions the CUDA kernel implements."""
    x = np.asarray(x, dtype=np.float32)
    n = len(x)
    f = np.float32
    if n < 8:
        r = f(0)
        for v in x:
            r = f(r + v)
        return r
    if n <= 128:
        r = [x[k] for k in range(8)]
        i = 8
        while i < n - (n % 8):
            for k in range(8):
                r[k] = f(r[k] + x[i + k])
            i += 8
        res = f(f(f(r[0] + r[1]) + f(r[2] + r[3])) + f(f(r[4] + r[5]) + f(r[6] + r[7])))
        while i < n:
            res = f(res + x[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return f(pairwise_sum_f32(x[:n2]) + pairwise_sum_f32(x[n2:]))


def color_run(results, data_dir, ext=".png"):
    """Matcher.run()'s colour branch (Matcher.py:328-336): results = sorted [(chi2, filename)]."""
    total_chi = sum([r[0] for r in results])
    total_matches = 300000. / total_chi
    raw = [(data_dir + "/" + r[1], 200. / r[0]) for r in results]
    total_prob = sum([x[1] for x in raw])
    raw_matches = [(x[0], x[1] / total_prob * total_matches) for x in raw]
    matches = sorted(raw_matches, key=lambda x: int(x[0].replace(ext, "").replace(data_dir + "/angle", "")))
    return total_matches, [x[1] / total_matches for x in matches]


# --------------------------------------------------------------------------------------------
# Laplacian variance (analyze.py:441-445)
# --------------------------------------------------------------------------------------------
def laplacian_int(img):
    """cv2.Laplacian(img, CV_64F) with ksize=1 is the 4-neighbour kernel [0 1 0;1 -4 1;0 1 0] with BORDER_REFLECT_101:
    integer valued in [-1020, 1020]."""
    a = np.asarray(img)
    assert a.ndim == 2
    p = np.pad(a.astype(np.int64), 1, mode="reflect")
    return p[:-2, 1:-1] + p[2:, 1:-1] + p[1:-1, :-2] + p[1:-1, 2:] - 4 * p[1:-1, 1:-1]


def pairwise_sum_f64(x):
    """numpy's pairwise summation of a contiguous float64 vector spelled out (numpy/_core/src/umath/loops_utils.h.src,
    DOUBLE_pairwise_sum): halve recursively (left part rounded down to a multiple of 8) until <= 128 elements; a piece is
    summed with 8 interleaved accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), leftovers one by one.
    Iterative (explicit stack) so that 480 000-element images do not recurse in Python; the accumulators are vectorised."""
    x = np.ascontiguousarray(x, dtype=np.float64)

    def leaf(lo, n):
        v = x[lo:lo + n]
        if n < 8:
            r = np.float64(0.0)
            for e in v:
                r = r + e
            return r
        m = n - n % 8
        r = v[:8].copy()
        for i in range(8, m, 8):
            r = r + v[i:i + 8]
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        for e in v[m:]:
            res = res + e
        return res

    def rec(lo, n):
        if n <= 128:
            return leaf(lo, n)
        n2 = n // 2
        n2 -= n2 % 8
        return rec(lo, n2) + rec(lo + n2, n - n2)   # depth is log2(n / 128): fine for any image

    return rec(0, len(x))


def laplacian_var(img):
    """analyze.py:441-445: cv2.Laplacian(img, cv2.CV_64F).var().  ndarray.var() is numpy's TWO-PASS variance,
    mean = sum(x) / N then sum((x - mean)^2) / N, both sums pairwise (pairwise_sum_f64).  The Laplacian is integer valued, so
    the first sum is exact; the second rounds at every addition and must be reproduced addition for addition, because the
    blur weight blur/200*0.85 (analyze.py:245-248) carries every bit of it into out.txt."""
    lap = laplacian_int(img)
    x = lap.astype(np.float64).ravel()
    n = np.float64(x.size)
    mean = np.float64(int(lap.sum())) / n
    d = x - mean
    return pairwise_sum_f64(d * d) / n


def laplacian_var_from_sums(img):
    """(sum L^2 - (sum L)^2 / N) / N on exact integer sums: within 1 ulp or so of laplacian_var, NOT bit-identical to it."""
    s1, s2, n = laplacian_sums(img)
    return (s2 - s1 * s1 / n) / n


def laplacian_sums(img):
    """(sum L, sum L^2, N) as exact Python ints: what the CUDA kernel must reproduce bit-exactly."""
    lap = laplacian_int(img)
    return int(lap.sum()), int((lap * lap).sum()), int(lap.size)


# --------------------------------------------------------------------------------------------
# particle filter update (analyze.py:239-336), generalised from 7x25 to L x I
# --------------------------------------------------------------------------------------------
def forward_factor(best_angle_index, step=15):
    return 0.05 * abs(math.sin(best_angle_index * step * 180 / math.pi))  # analyze.py:331 (sic); step = 15 in the reference


def account_command(command, previous, step=15):
    """analyze.py:312-336.  ``previous`` is a list of [total, [p...]]; like the reference this
    mutates the inner lists of ``previous`` (shallow copy at :317)."""
    copy = previous[:]
    n_loc = len(previous)
    if command == "l":
        for c in copy:
            c[1] = c[1][1:] + c[1][0:1]
    elif command == "r":
        for c in copy:
            c[1] = c[1][-1:] + c[1][0:-1]
    elif command == "f":
        bc = previous.index(max(previous))
        ba = previous[bc][1].index(max(previous[bc][1]))
        factor = forward_factor(ba, step)
        if bc < n_loc - 1 and ba * step < 180 and ba > 0:
            copy[bc + 1][0] *= (1 + factor)
        elif bc > 0 and ba * step > 180 and ba * step < 360:
            copy[bc - 1][0] *= (1 + factor)
    return copy


def _blend(previous, current, cw, pw):
    out = []
    for p, c in zip(previous, current):
        out.append([cw * c[0] + pw * p[0], [cw * ci + pw * pi for ci, pi in zip(c[1], p[1])]])
    return out


def prev_weight(previous, current):
    cw = 0.7
    return _blend(previous, current, cw, 1 - cw)  # analyze.py:282-283


def prob_update(previous, current, blur):
    cw = 0.85 if blur > 200 else (blur / 200) * 0.85  # analyze.py:245-248
    return _blend(previous, current, cw, 1 - cw)


def best_guess(adjusted):
    """analyze.py:133-134: lexicographic max over [total, probs] lists, first argmax of probs."""
    bc = adjusted.index(max(adjusted))
    ba = adjusted[bc][1].index(max(adjusted[bc][1]))
    return bc, ba


def filter_step(command, previous, current, blur, step=15):
    """One frame of processRaw (analyze.py:119-138).  Returns (adjusted, (bestCircle, bestAngle))."""
    acc = account_command(command, previous, step)
    adj = prev_weight(acc, current)
    adj = prob_update(acc, adj, blur)
    return adj, best_guess(adj)


def initial_probs(n_loc=7, n_img=25):
    return [[1, [1 / (3 * n_img)] * n_img] for _ in range(n_loc)]  # analyze.py:106-107: [1/75]*25, i.e. 1/(3 I)


# --------------------------------------------------------------------------------------------
# text formats (analyze.py:351-398)
# --------------------------------------------------------------------------------------------
def write_prob(prob, filename):
    # legacy print options: np.float32 list elements (colour method) print as plain numbers, the format the golden run
    # used (ref_shims item 4) and the only one analyzer.readProb can parse back
    with open(filename, "w") as f, np.printoptions(legacy="1.25"):
        for item in prob:
            f.write(str(item[0]) + "\n")
            f.write(str(item[1]) + "\n")


def read_prob(filename, n_loc=7):
    with open(filename) as f:
        raw = f.read().split("\n")[:-1]
    chunks = [raw[i:i + 2] for i in range(0, len(raw), 2)]
    frames = [chunks[i:i + n_loc] for i in range(0, len(chunks), n_loc)]
    out = {}
    for k, fr in enumerate(frames):
        out[str(k).zfill(4)] = [
            [float(loc[0]), list(map(float, loc[1].replace("[", "").replace("]", "").split(",")))]
            for loc in fr
        ]
    return out


# --------------------------------------------------------------------------------------------
# cv2-backed versions: the third-party routines the reference actually calls.  Used to pin the
# restatements above and as the CPU baseline ("B - oracle matching only" in BASELINE.md).
# --------------------------------------------------------------------------------------------
def cv2_l2_pair_count(des1, des2, ratio=RATIO):
    import cv2
    if des1 is None or des2 is None or len(des1) == 0 or len(des2) < 2:
        return 0
    bf = cv2.BFMatcher(cv2.NORM_L2)
    matches = bf.knnMatch(np.ascontiguousarray(des1, dtype=np.float32),
                          np.ascontiguousarray(des2, dtype=np.float32), k=2)
    filtered = list(filter(lambda x: x[0].distance < ratio * x[1].distance, matches))
    return len(filtered)


def cv2_orb_pair_count_knn(des1, des2, ratio=RATIO):
    import cv2
    if des1 is None or des2 is None or len(des1) == 0 or len(des2) < 2:
        return 0
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    matches = bf.knnMatch(np.ascontiguousarray(des1), np.ascontiguousarray(des2), k=2)
    return len(list(filter(lambda x: x[0].distance < ratio * x[1].distance, matches)))


def cv2_orb_pair_count_crosscheck(des1, des2):
    import cv2
    if des1 is None or des2 is None or len(des1) == 0 or len(des2) == 0:
        return 0
    bf = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True)
    return len(bf.match(np.ascontiguousarray(des1), np.ascontiguousarray(des2)))


def cv2_laplacian_var(img):
    import cv2
    return cv2.Laplacian(np.ascontiguousarray(img), cv2.CV_64F).var()


def scipy_bow_score(des, voc, idf, im_features, num_words):
    """Matcher.py:249-258 verbatim with the third-party calls."""
    from scipy.cluster.vq import vq
    from sklearn import preprocessing
    test_features = np.zeros((1, num_words), "float32")
    words, distance = vq(des, voc)
    for w in words:
        test_features[0][w] += 1
    test_features = test_features * idf
    test_features = preprocessing.normalize(test_features, norm="l2")
    return words, np.dot(test_features, np.asarray(im_features).T)
