"""Seeded synthetic inputs of the shapes BASELINE.md names (no datasets ship with the reference).

Descriptor-level generators return plain numpy arrays laid out like cv2 extractor output
(SIFT: float32 integer-valued 0..255 x128; ORB: uint8 x32; SURF: float32 unit-norm x64).
Image-level generators draw a wide "panorama" strip of random shapes and crop one view per
angle so that neighbouring views genuinely match (map/<l>/angleNNN.png, cam1_img/NNNN.png,
commands.txt as analyze.py:42,52,78 and Matcher.py:310 expect).
"""
from __future__ import annotations

import os

import numpy as np


# ---------------------------------------------------------------------------------------------
# descriptor level
# ---------------------------------------------------------------------------------------------
def sift_like(rng, n, dim=128):
    """Integer-valued SIFT-like rows: round(512*g/|g|), g ~ Gamma(0.6), clipped to 0..255."""
    g = rng.gamma(0.6, 1.0, size=(n, dim))
    g /= np.maximum(np.linalg.norm(g, axis=1, keepdims=True), 1e-12)
    return np.clip(np.rint(512.0 * g), 0, 255).astype(np.float32)


def sift_noisy_copy(rng, rows, sigma=6.0):
    """A re-observation of the same keypoints: additive integer noise, clipped to 0..255."""
    noise = np.rint(rng.normal(0.0, sigma, size=rows.shape))
    return np.clip(rows + noise, 0, 255).astype(np.float32)


def sift_map_and_frames(seed, n_images, nm, n_frames, nq, planted_frac=0.4, sigma=6.0):
    """Map of n_images x (nm,128) and n_frames queries of (nq,128).  Frame f re-observes
    `planted_frac` of the rows of map image (f*7) % n_images, so exactly one map image per frame
    carries many ratio-test survivors and the rest carry ~none (like a real localisation run)."""
    rng = np.random.default_rng(seed)
    map_des = [sift_like(rng, nm) for _ in range(n_images)]
    frames = []
    n_pl = int(round(planted_frac * min(nq, nm)))
    for f in range(n_frames):
        q = sift_like(rng, nq)
        src = map_des[(f * 7) % n_images]
        if n_pl:
            rows = rng.choice(nm, size=n_pl, replace=False)
            dst = rng.choice(nq, size=n_pl, replace=False)
            q[dst] = sift_noisy_copy(rng, src[rows], sigma)
        frames.append(q)
    return map_des, frames


def orb_like(rng, n):
    return rng.integers(0, 256, size=(n, 32), dtype=np.uint8)


def orb_flip(rng, rows, max_flips=40):
    """Bit-flipped (0..max_flips flips per row) copies."""
    out = rows.copy()
    for r in range(out.shape[0]):
        k = int(rng.integers(0, max_flips + 1))
        if k:
            bits = rng.choice(256, size=k, replace=False)
            for b in bits:
                out[r, b >> 3] ^= np.uint8(1 << (b & 7))
    return out


def orb_map_and_query(seed, n_images=50, nm=500, nq=500, planted_frac=0.4):
    """BASELINE cfg1: uniform random bytes with planted_frac of the query rows being bit-flipped
    copies of map rows (spread over the images)."""
    rng = np.random.default_rng(seed)
    map_des = [orb_like(rng, nm) for _ in range(n_images)]
    q = orb_like(rng, nq)
    n_pl = int(round(planted_frac * nq))
    dst = rng.choice(nq, size=n_pl, replace=False)
    for d in dst:
        img = int(rng.integers(0, n_images))
        row = int(rng.integers(0, nm))
        q[d] = orb_flip(rng, map_des[img][row:row + 1])[0]
    return map_des, q


def surf_like(rng, n, dim=64):
    """64-d unit-norm signed float32 rows (Gaussian x Gamma magnitudes, normalised)."""
    g = rng.normal(size=(n, dim)) * rng.gamma(0.8, 1.0, size=(n, dim))
    g /= np.maximum(np.linalg.norm(g, axis=1, keepdims=True), 1e-12)
    return g.astype(np.float32)


def surf_noisy_copy(rng, rows, sigma=0.02):
    g = rows + rng.normal(0.0, sigma, size=rows.shape)
    g /= np.maximum(np.linalg.norm(g, axis=1, keepdims=True), 1e-12)
    return g.astype(np.float32)


def bow_index(seed, n_images, num_words, voc_rows=None, des_per_image=300, iters=4):
    """A per-location BOW index with the reference's layout (Matcher.py:124-141):
    (im_features (N,W) f32 L2-normalised tf-idf, idf (W,) f32, voc (<=W,128) f32).
    The codebook is a few Lloyd iterations from a seeded sample (scipy's kmeans is unseeded, and
    only the *query* side, Matcher.py:232-259, is on the hot path)."""
    rng = np.random.default_rng(seed)
    voc_rows = voc_rows or num_words
    des = [sift_like(rng, des_per_image) for _ in range(n_images)]
    allv = np.vstack(des)
    voc = allv[rng.choice(len(allv), size=min(voc_rows, len(allv)), replace=False)].astype(np.float64)
    voc += rng.normal(0, 0.25, size=voc.shape)  # fp32, non-integer centroids as kmeans produces
    for _ in range(iters):
        lab = _assign(allv, voc)
        for j in np.unique(lab):
            voc[j] = allv[lab == j].mean(0)
    voc = voc.astype(np.float32)
    im = np.zeros((n_images, num_words), "float32")
    for i, d in enumerate(des):
        for w in _assign(d, voc.astype(np.float64)):
            im[i][w] += 1
    occ = np.sum((im > 0) * 1, axis=0)
    idf = np.array(np.log((1.0 * n_images + 1) / (1.0 * occ + 1)), "float32")
    im = im * idf
    nrm = np.sqrt((im.astype(np.float64) ** 2).sum(1, keepdims=True))
    nrm[nrm == 0] = 1
    im = (im / nrm).astype(np.float32)
    return im, idf, voc, des


def _assign(o, c, block=4096):
    o = np.asarray(o, np.float64)
    cn = (c * c).sum(1)[None, :]
    out = np.empty(len(o), np.int64)
    for s in range(0, len(o), block):
        ob = o[s:s + block]
        out[s:s + block] = (cn - 2.0 * ob @ c.T).argmin(1)
    return out


# ---------------------------------------------------------------------------------------------
# image level
# ---------------------------------------------------------------------------------------------
def _draw_strip(rng, height, width, n_shapes):
    import cv2
    img = np.full((height, width, 3), 127, np.uint8)
    img[:] = rng.integers(60, 200, size=3, dtype=np.uint8)
    for _ in range(n_shapes):
        kind = int(rng.integers(0, 3))
        col = tuple(int(c) for c in rng.integers(0, 256, size=3))
        x, y = int(rng.integers(0, width)), int(rng.integers(0, height))
        if kind == 0:
            w, h = int(rng.integers(6, height // 2)), int(rng.integers(6, height // 2))
            cv2.rectangle(img, (x, y), (x + w, y + h), col, -1 if rng.random() < 0.6 else 2)
        elif kind == 1:
            cv2.circle(img, (x, y), int(rng.integers(4, height // 4)), col, -1 if rng.random() < 0.6 else 2)
        else:
            x2, y2 = int(rng.integers(0, width)), int(rng.integers(0, height))
            if abs(x2 - x) > width // 6:
                x2 = x + int(np.sign(x2 - x)) * width // 6
            cv2.line(img, (x, y), (x2, y2), col, int(rng.integers(1, 4)))
    tex = rng.integers(-12, 13, size=(height, width, 1))
    return np.clip(img.astype(np.int16) + tex, 0, 255).astype(np.uint8)


def make_dataset(root, seed=0, n_locations=7, n_images=25, n_frames=6, view=(160, 120),
                 n_shapes=260, commands="lrfbs", angle_step=15):
    """Writes map/<l>/angleNNN.png, cam1_img/NNNN.png and commands.txt under `root`."""
    import cv2
    rng = np.random.default_rng(seed)
    vw, vh = view
    step = vw // 3
    strips = []
    for loc in range(n_locations):
        strip = _draw_strip(rng, vh, step * n_images + vw, n_shapes)
        strips.append(strip)
        d = os.path.join(root, "map", str(loc))
        os.makedirs(d, exist_ok=True)
        for k in range(n_images):
            crop = strip[:, k * step:k * step + vw]
            cv2.imwrite(os.path.join(d, "angle" + str(k * angle_step).zfill(3) + ".png"), crop)
    os.makedirs(os.path.join(root, "cam1_img"), exist_ok=True)
    lines = []
    for f in range(n_frames):
        loc = (f // 2) % n_locations
        k = (3 * f + 1) % n_images
        off = int(rng.integers(-step // 4, step // 4 + 1))
        x0 = int(np.clip(k * step + off, 0, strips[loc].shape[1] - vw))
        crop = strips[loc][:, x0:x0 + vw].astype(np.int16)
        crop = np.clip(crop + rng.integers(-6, 7, size=crop.shape), 0, 255).astype(np.uint8)
        if f % 3 == 2:
            crop = cv2.GaussianBlur(crop, (0, 0), 1.2)
        cv2.imwrite(os.path.join(root, "cam1_img", str(f).zfill(4) + ".png"), crop)
        lines.append(str(f).zfill(4) + " " + commands[f % len(commands)])
    with open(os.path.join(root, "commands.txt"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    return root


def textured_views(seed, n_views=48, n_queries=12, view=(800, 600), shapes_per_view=600):
    """Overlapping views of one richly textured synthetic panorama (map images) and re-observations of some of them with
    pixel noise and a small shift (query frames), as BGR uint8 arrays.  cv2.SIFT finds 1-3 k keypoints per 800x600 view: the
    image-level stand-in for the README's 800x600 case (no datasets ship with the reference)."""
    rng = np.random.default_rng(seed)
    vw, vh = view
    step = vw // 3
    width = step * n_views + vw
    strip = _draw_strip(rng, vh, width, int(shapes_per_view * width / vw))
    views = [np.ascontiguousarray(strip[:, k * step:k * step + vw]) for k in range(n_views)]
    queries = []
    for f in range(n_queries):
        k = (5 * f + 2) % n_views
        off = int(rng.integers(-step // 5, step // 5 + 1))
        x0 = int(np.clip(k * step + off, 0, strip.shape[1] - vw))
        crop = strip[:, x0:x0 + vw].astype(np.int16)
        queries.append(np.clip(crop + rng.integers(-5, 6, size=crop.shape), 0, 255).astype(np.uint8))
    return views, queries
