"""GPU parity tests: the CUDA path (through the C-ABI, via ctypes) against the CPU oracle on the same seeded
inputs, and against the golden fixtures produced by the unmodified reference (tests/golden/, oracle/make_golden.py).
Bit-exact for SIFT / ORB counts, Laplacian sums and the filter update; SURF / BOW under the stated tolerances."""
import json
import os

import numpy as np
import pytest

import oracle  # oracle/oracle.py: test infrastructure only
import pymcl_b200  # noqa: F401
from pymcl_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def split(mat, off):
    return [mat[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def oracle_counts(fn, qs, ms):
    return np.array([[fn(q, m) for m in ms] for q in qs], dtype=np.int32)


def tensor_launch_variants(ctx):
    """The tensor-core kernels launch as plain CTAs or as CTA pairs sharing the map tiles by multicast (chosen from the batch
    shape): yields after forcing each (option 5: 1 = never pairs, 2 = always pairs, 0 = automatic)."""
    try:
        for cl in (1, 2, 0):
            ctx.set_option(5, cl)
            yield cl
    finally:
        ctx.set_option(5, 0)


# ------------------------------------------------------------------------------------------------
# SIFT
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("algo", [_lib.ALGO_SIMT, _lib.ALGO_TENSOR, _lib.ALGO_AUTO])
def test_sift_golden_reference_counts(ctx, golden_dir, algo):
    """Descriptors extracted by the reference run + the reference's own SIFTMatch counts (exact search)."""
    z = np.load(os.path.join(golden_dir, "sift_desc_case.npz"))
    qs, ms = split(z["q"], z["q_off"]), split(z["m"], z["m_off"])
    m = _lib.MapIndex(ctx, "SIFT", ms)
    got = m.match_counts(qs, algo=algo)
    assert np.array_equal(got, z["counts"])
    # float32 integer-valued input (what cv2 hands over) takes the same path
    got32 = m.match_counts([q.astype(np.float32) for q in qs], algo=algo)
    assert np.array_equal(got32, z["counts"])
    m.close()


@pytest.mark.parametrize("algo", [_lib.ALGO_SIMT, _lib.ALGO_TENSOR])
@pytest.mark.parametrize("nm,nq,n_images,n_frames", [(256, 256, 12, 5), (300, 150, 9, 3), (37, 61, 20, 4),
                                                     (1000, 700, 3, 2)])
def test_sift_synthetic_vs_oracle(ctx, algo, nm, nq, n_images, n_frames):
    map_des, frames = synth.sift_map_and_frames(11 + nm, n_images, nm, n_frames, nq)
    want = oracle_counts(oracle.sift_pair_count, frames, map_des)
    m = _lib.MapIndex(ctx, "SIFT", map_des)
    got = m.match_counts(frames, algo=algo)
    m.close()
    assert want.sum() > 0
    assert np.array_equal(got, want)


@pytest.mark.parametrize("algo", [_lib.ALGO_SIMT, _lib.ALGO_TENSOR])
def test_sift_ragged_and_edge_cases(ctx, algo):
    """Variable descriptors per image incl. 0, 1, 2, 15, 16, 17, 128, 129 rows; empty frames; duplicates."""
    rng = np.random.default_rng(5)
    sizes = [0, 1, 2, 15, 16, 17, 128, 129, 0, 255, 3, 640]
    map_des = [synth.sift_like(rng, n) if n else None for n in sizes]
    map_des[4][3] = map_des[4][9]  # duplicate rows inside one image: d0 == d1 -> never a match
    frames = [synth.sift_like(rng, 200), None, synth.sift_like(rng, 1), synth.sift_like(rng, 513)]
    frames[0][:50] = synth.sift_noisy_copy(rng, map_des[11][:50], 4.0)
    frames[3][:100] = synth.sift_noisy_copy(rng, map_des[9][:100], 3.0)
    frames[3][100:110] = map_des[3][:10]  # exact copies (d0 = 0)
    want = oracle_counts(oracle.sift_pair_count, frames, map_des)
    m = _lib.MapIndex(ctx, "SIFT", map_des)
    got = m.match_counts(frames, algo=algo)
    m.close()
    assert np.array_equal(got, want)
    assert got[1].sum() == 0 and got[:, 0].sum() == 0 and got[:, 1].sum() == 0


@pytest.mark.parametrize("algo", [_lib.ALGO_TENSOR])
def test_sift_many_tiny_images_and_one_huge_image(ctx, algo):
    """Shapes at the edges of the tiling: hundreds of images of a few rows each (every 32-row group is another image) and
    one image longer than the shared-memory metadata window of a chunk (> 352 tiles of 64 rows)."""
    rng = np.random.default_rng(31)
    tiny = [synth.sift_like(rng, int(n)) for n in rng.integers(2, 40, size=300)]
    frames = [synth.sift_like(rng, 500), synth.sift_like(rng, 130)]
    src = np.vstack(tiny[:40])
    frames[0][:200] = synth.sift_noisy_copy(rng, src[:200], 3.0)
    want = oracle_counts(oracle.sift_pair_count, frames, tiny)
    m = _lib.MapIndex(ctx, "SIFT", tiny)
    got = m.match_counts(frames, algo=algo)
    m.close()
    assert want.sum() > 100
    assert np.array_equal(got, want)
    huge = [synth.sift_like(rng, 23000), synth.sift_like(rng, 70), synth.sift_like(rng, 2050)]
    frames[1][:60] = synth.sift_noisy_copy(rng, huge[0][22000:22060], 3.0)
    want = oracle_counts(oracle.sift_pair_count, frames, huge)
    m = _lib.MapIndex(ctx, "SIFT", huge)
    got = m.match_counts(frames, algo=algo)
    m.close()
    assert want[1, 0] >= 50
    assert np.array_equal(got, want)


def test_sift_adversarial_norm_spread_forces_exact_rescans(ctx):
    """Rows whose norms vary by orders of magnitude inside every tile (so the tile brackets are useless) and many
    near-threshold best/second pairs: the bracketed pre-filter cannot decide, the fix-up kernel must fall back to exact
    whole-image rescans -- and the counts must still equal the exact SIMT kernel's and the oracle's."""
    rng = np.random.default_rng(41)

    def wild(n):
        base = synth.sift_like(rng, n)
        scale = rng.choice([0.05, 0.2, 0.5, 1.0], size=(n, 1))
        return np.clip(np.rint(base * scale), 0, 255).astype(np.float32)

    map_des = [wild(n) for n in (700, 90, 333, 64, 1500)]
    frames = [wild(600), wild(257)]
    # near-threshold structure: query rows that are convex mixtures of two map rows
    for f, (i, n) in enumerate(((0, 300), (4, 200))):
        a = map_des[i][rng.integers(0, len(map_des[i]), size=n)]
        b = map_des[i][rng.integers(0, len(map_des[i]), size=n)]
        w = rng.uniform(0.55, 0.75, size=(n, 1))
        frames[f][:n] = np.clip(np.rint(w * a + (1 - w) * b), 0, 255)
    before = ctx.info()["sift_rescans"]
    m = _lib.MapIndex(ctx, "SIFT", map_des)
    simt = m.match_counts(frames, algo=_lib.ALGO_SIMT)
    for cl in tensor_launch_variants(ctx):
        assert np.array_equal(m.match_counts(frames, algo=_lib.ALGO_TENSOR), simt), cl
    m.close()
    assert np.array_equal(simt, oracle_counts(oracle.sift_pair_count, frames, map_des))
    assert simt.sum() > 50
    assert ctx.info()["sift_rescans"] > before, "the exact-rescan path was not exercised"


@pytest.mark.parametrize("seed", range(12))
def test_sift_fuzz_all_generations_equal_exact_kernel(ctx, seed):
    """Random shapes and data styles (SIFT-like, wildly scaled norms, low-entropy rows with many exact ties, copies,
    mixtures near the ratio threshold): every tensor-core generation must equal the exact SIMT kernel, which itself is
    checked against the oracle on a sample."""
    rng = np.random.default_rng(1000 + seed)

    def rows(n, style):
        if style == 0:
            return synth.sift_like(rng, n)
        if style == 1:
            return np.clip(np.rint(synth.sift_like(rng, n) * rng.choice([0.03, 0.1, 0.4, 1.0], size=(n, 1))), 0, 255).astype(np.float32)
        if style == 2:
            return (rng.integers(0, 3, size=(n, 128)) * rng.choice([1, 60, 127], size=(n, 1))).astype(np.float32)
        return rng.integers(0, 256, size=(n, 128)).astype(np.float32)

    n_img = int(rng.integers(1, 9))
    map_des = [rows(int(rng.integers(0, 900)), int(rng.integers(0, 4))) for _ in range(n_img)]
    map_des = [m if len(m) else None for m in map_des]
    frames = [rows(int(rng.integers(1, 700)), int(rng.integers(0, 4))) for _ in range(int(rng.integers(1, 4)))]
    for f in frames:   # plant copies / noisy copies / mixtures
        src = [m for m in map_des if m is not None and len(m) >= 2]
        if not src:
            break
        m = src[int(rng.integers(0, len(src)))]
        n = min(len(f), len(m)) // 2
        if n:
            a, b = m[rng.integers(0, len(m), size=n)], m[rng.integers(0, len(m), size=n)]
            w = rng.choice([1.0, 0.9, 0.7, 0.6, 0.5], size=(n, 1))
            f[:n] = np.clip(np.rint(w * a + (1 - w) * b + rng.normal(0, rng.choice([0, 1, 4]), size=a.shape)), 0, 255)
    mi = _lib.MapIndex(ctx, "SIFT", map_des)
    exact = mi.match_counts(frames, algo=_lib.ALGO_SIMT)
    for cl in tensor_launch_variants(ctx):
        assert np.array_equal(mi.match_counts(frames, algo=_lib.ALGO_TENSOR), exact), (seed, cl)
    mi.close()
    f0 = frames[0]
    assert np.array_equal(exact[0], [oracle.sift_pair_count(f0, m) for m in map_des])


def test_sift_candidate_overflow_is_repaired_on_the_device(ctx):
    """A candidate that does not fit the list marks its (frame, image) pair dirty; the exact SIMT kernel then recomputes
    exactly those pairs in the same call (no host round trip, no error): results are unchanged."""
    map_des, frames = synth.sift_map_and_frames(5, 6, 400, 3, 300)
    m = _lib.MapIndex(ctx, "SIFT", map_des)
    want = m.match_counts(frames, algo=_lib.ALGO_SIMT)
    assert want.sum() > 200
    repaired0 = ctx.stats()["repaired_pairs"]
    ctx.set_option(3, 16)
    try:
        for cl in tensor_launch_variants(ctx):
            assert np.array_equal(m.match_counts(frames, algo=_lib.ALGO_TENSOR), want), cl
        assert ctx.stats()["repaired_pairs"] > repaired0, "the dirty-pair repair did not run"
    finally:
        ctx.set_option(3, 0)
    m.close()


def test_sift_tile_records_from_global_memory_path(ctx):
    """sift_tc3 / sift_tc4 stage a chunk's 64-row records in shared memory; chunks longer than the staging area (a single
    image of > 65 536 rows for tc4) read them from global memory instead.  Option 1 forces that path on ordinary sizes."""
    map_des, frames = synth.sift_map_and_frames(21, 9, 700, 4, 500)
    m = _lib.MapIndex(ctx, "SIFT", map_des)
    want = m.match_counts(frames, algo=_lib.ALGO_SIMT)
    assert want.sum() > 100
    ctx.set_option(1, 1)
    try:
        for cl in tensor_launch_variants(ctx):
            assert np.array_equal(m.match_counts(frames, algo=_lib.ALGO_TENSOR), want), cl
    finally:
        ctx.set_option(1, 0)
    m.close()


def test_sift_ratio_known_answers(ctx, golden_dir):
    """The cv2 ratio semantics float(sqrtf(d0^2)) < 0.7*float(sqrtf(d1^2)) on the survey's integer pairs:
    build 1 query row and 2 map rows with exactly those squared distances."""
    kat = json.load(open(os.path.join(golden_dir, "kat.json")))["sift_ratio"]

    def rows_with_d2(d2):
        # decompose d2 into a sum of squares <= 255^2 spread over 128 dims
        v = np.zeros(128, dtype=np.int64)
        rem = d2
        i = 0
        while rem > 0:
            s = min(255, int(np.floor(np.sqrt(rem))))
            v[i] = s
            rem -= s * s
            i += 1
        assert i <= 128
        return v

    for algo in (_lib.ALGO_SIMT, _lib.ALGO_TENSOR):
        for d0, d1, expect in kat:
            q = np.zeros((1, 128), np.uint8)
            m = np.stack([rows_with_d2(d0), rows_with_d2(d1)]).astype(np.uint8)
            assert oracle.sift_pair_count(q, m) == int(expect)
            mi = _lib.MapIndex(ctx, "SIFT", [m])
            got = mi.match_counts([q], algo=algo)
            mi.close()
            assert int(got[0, 0]) == int(expect), (d0, d1, algo)


def test_sift_mask_fill(ctx):
    map_des, frames = synth.sift_map_and_frames(3, 10, 120, 4, 90)
    want = oracle_counts(oracle.sift_pair_count, frames, map_des)
    rng = np.random.default_rng(0)
    mask = (rng.random((4, 10)) < 0.4).astype(np.uint8)
    m = _lib.MapIndex(ctx, "SIFT", map_des)
    for algo in (_lib.ALGO_AUTO, _lib.ALGO_TENSOR, _lib.ALGO_SIMT):
        got = m.match_counts(frames, mask=mask, fill_value=1, algo=algo)
        assert np.array_equal(got, np.where(mask > 0, want, 1))
    m.close()


def test_sift_rejects_non_integer(ctx):
    bad = np.full((4, 128), 0.5, dtype=np.float32)
    with pytest.raises(_lib.MclError):
        _lib.MapIndex(ctx, "SIFT", [bad])


def test_sift_full_size_properties(ctx):
    """BASELINE cfg3 tile shape (2048 x 2048 per pair) at a reduced image count: size-independent properties.
    (a) planted copies are found: the planted image gets >= 90% of its planted rows, others stay near 0;
    (b) permuting the map images permutes the counts; (c) tensor-core and SIMT paths agree bit-exactly."""
    map_des, frames = synth.sift_map_and_frames(3, 6, 2048, 2, 2048, planted_frac=0.25, sigma=4.0)
    m = _lib.MapIndex(ctx, "SIFT", map_des)
    tc = m.match_counts(frames, algo=_lib.ALGO_TENSOR)
    simt = m.match_counts(frames, algo=_lib.ALGO_SIMT)
    m.close()
    assert np.array_equal(tc, simt)
    for f in range(2):
        planted = (f * 7) % 6
        assert tc[f, planted] >= 0.9 * 512
        assert np.delete(tc[f], planted).max() < 64
    perm = [3, 0, 5, 1, 4, 2]
    m2 = _lib.MapIndex(ctx, "SIFT", [map_des[i] for i in perm])
    tc2 = m2.match_counts(frames, algo=_lib.ALGO_TENSOR)
    m2.close()
    assert np.array_equal(tc2, tc[:, perm])


def test_sift_host_batch_pipelined_equals_resident(ctx):
    """Host batches above 64 k rows are cut into parts (H2D + ingest of part p+1 overlaps the matching of part p):
    same counts as the resident-query path and as the oracle on a sample; ragged frames, a mask, float32 input."""
    import torch
    rng = np.random.default_rng(12)
    map_des = [synth.sift_like(rng, n) for n in (700, 64, 1500, 333)]
    sizes = [2048] * 30 + [0, 1, 777, 4000, 2048, 2048, 5, 2048, 3000, 2048]
    frames = [synth.sift_like(rng, n) if n else None for n in sizes]
    frames[3][:300] = synth.sift_noisy_copy(rng, map_des[2][:300], 4.0)
    frames[33][:200] = synth.sift_noisy_copy(rng, map_des[0][:200], 4.0)
    assert sum(sizes) > 2 * 32768
    m = _lib.MapIndex(ctx, "SIFT", map_des)
    got = m.match_counts(frames, algo=_lib.ALGO_TENSOR)
    q = _lib.QueryBatch(ctx, "SIFT", frames)
    out = torch.zeros((len(frames), len(map_des)), dtype=torch.int32, device="cuda")
    m.match_counts_device(q, out.data_ptr(), algo=_lib.ALGO_TENSOR)
    ctx.sync()
    assert np.array_equal(got, out.cpu().numpy())
    for f in (3, 30, 31, 32, 33, 39):
        assert np.array_equal(got[f], [oracle.sift_pair_count(frames[f], d) for d in map_des])
    assert got[3, 2] >= 250 and got[33, 0] >= 150
    mask = (rng.random(got.shape) < 0.5).astype(np.uint8)
    assert np.array_equal(m.match_counts(frames, mask=mask, fill_value=1, algo=_lib.ALGO_TENSOR), np.where(mask > 0, got, 1))
    bad = [f.copy() if f is not None else None for f in frames]
    bad[20][5, 5] = 0.25
    with pytest.raises(_lib.MclError):
        m.match_counts(bad, algo=_lib.ALGO_TENSOR)
    q.close()
    m.close()


# ------------------------------------------------------------------------------------------------
# ORB
# ------------------------------------------------------------------------------------------------
ORB_ALGOS = [_lib.ALGO_SIMT, _lib.ALGO_TENSOR, _lib.ALGO_AUTO]


@pytest.mark.parametrize("algo", ORB_ALGOS)
def test_orb_golden_reference_counts(ctx, golden_dir, algo):
    z = np.load(os.path.join(golden_dir, "orb_desc_case.npz"))
    qs, ms = split(z["q"], z["q_off"]), split(z["m"], z["m_off"])
    m = _lib.MapIndex(ctx, "ORB", ms)
    assert np.array_equal(m.match_counts(qs, mode=_lib.KNN2_RATIO, algo=algo), z["counts_knn"])
    assert np.array_equal(m.match_counts(qs, mode=_lib.CROSSCHECK, algo=algo), z["counts_crosscheck"])
    m.close()


@pytest.mark.parametrize("algo", ORB_ALGOS)
def test_orb_cfg1_vs_oracle(ctx, algo):
    """BASELINE config 1 at full size: one (500,32) query vs 50 images of (500,32)."""
    map_des, q = synth.orb_map_and_query(1)
    m = _lib.MapIndex(ctx, "ORB", map_des)
    want_knn = oracle_counts(oracle.orb_pair_count_knn, [q], map_des)
    want_cc = oracle_counts(oracle.orb_pair_count_crosscheck, [q], map_des)
    assert want_knn.sum() > 50
    for cl in tensor_launch_variants(ctx):
        assert np.array_equal(m.match_counts([q], mode=_lib.KNN2_RATIO, algo=algo), want_knn), cl
    assert np.array_equal(m.match_counts([q], mode=_lib.CROSSCHECK, algo=algo), want_cc)
    m.close()


@pytest.mark.parametrize("algo", ORB_ALGOS)
def test_orb_ragged_ties_and_mask(ctx, algo):
    rng = np.random.default_rng(9)
    sizes = [0, 1, 2, 3, 127, 128, 129, 257, 700, 31, 32, 33, 5]
    map_des = [synth.orb_like(rng, n) if n else None for n in sizes]
    # low-entropy rows -> many distance ties (first-index tie-breaking matters for crossCheck)
    map_des[6] = (rng.integers(0, 2, size=(129, 32)) * 255).astype(np.uint8)
    # an image whose rows are all far from every query row (Hamming > 128): the zero rows that pad its last group must not win
    frames = [synth.orb_like(rng, 300), (rng.integers(0, 2, size=(140, 32)) * 255).astype(np.uint8), None,
              synth.orb_like(rng, 1), synth.orb_like(rng, 257)]
    map_des[12] = (~frames[4][:5]).copy()
    map_des[12][1, 0] ^= 3
    frames[0][:80] = synth.orb_flip(rng, map_des[8][:80], 30)
    frames[4][100:180] = synth.orb_flip(rng, map_des[8][300:380], 20)
    want_knn = oracle_counts(oracle.orb_pair_count_knn, frames, map_des)
    want_cc = oracle_counts(oracle.orb_pair_count_crosscheck, frames, map_des)
    m = _lib.MapIndex(ctx, "ORB", map_des)
    for cl in tensor_launch_variants(ctx):
        assert np.array_equal(m.match_counts(frames, mode=_lib.KNN2_RATIO, algo=algo), want_knn), cl
    assert np.array_equal(m.match_counts(frames, mode=_lib.CROSSCHECK, algo=algo), want_cc)
    mask = (rng.random(want_knn.shape) < 0.5).astype(np.uint8)
    assert np.array_equal(m.match_counts(frames, mode=_lib.KNN2_RATIO, mask=mask, fill_value=1, algo=algo),
                          np.where(mask > 0, want_knn, 1))
    assert np.array_equal(m.match_counts(frames, mode=_lib.CROSSCHECK, mask=mask, fill_value=7, algo=algo),
                          np.where(mask > 0, want_cc, 7))
    m.close()


@pytest.mark.parametrize("seed", range(8))
def test_orb_tensor_core_fuzz_equals_exact_kernel(ctx, seed):
    """ORB 2-NN + ratio on the tensor cores (+-1 rows, s8 x s8 -> s32, exact keys) against the xor + popc SIMT kernel and,
    on a sample, the oracle: random shapes (tiny images, images of thousands of rows), random / low-entropy / duplicated
    rows, bit-flipped copies at every distance, ratios other than 0.7."""
    rng = np.random.default_rng(500 + seed)

    def rows(n, style):
        if style == 0:
            return synth.orb_like(rng, n)
        if style == 1:
            return (rng.integers(0, 2, size=(n, 32)) * 255).astype(np.uint8)
        base = synth.orb_like(rng, max(1, n // 7 + 1))
        return synth.orb_flip(rng, base[rng.integers(0, len(base), size=n)], int(rng.integers(1, 60)))

    n_img = int(rng.integers(1, 40))
    big = int(rng.integers(0, n_img))
    map_des = [rows(int(rng.integers(0, 700)) if i != big else int(rng.integers(2000, 6000)), int(rng.integers(0, 3))) for i in range(n_img)]
    map_des = [m if len(m) else None for m in map_des]
    frames = [rows(int(rng.integers(1, 900)), int(rng.integers(0, 3))) for _ in range(int(rng.integers(1, 6)))]
    for f in frames:
        src = [m for m in map_des if m is not None and len(m) >= 2]
        m = src[int(rng.integers(0, len(src)))]
        n = min(len(f), len(m)) // 2
        if n:
            f[:n] = synth.orb_flip(rng, m[rng.integers(0, len(m), size=n)], int(rng.integers(0, 90)))
    ratio = float(rng.choice([0.7, 0.7, 0.5, 0.9, 1.0]))
    mi = _lib.MapIndex(ctx, "ORB", map_des)
    exact = mi.match_counts(frames, ratio=ratio, algo=_lib.ALGO_SIMT)
    for cl in tensor_launch_variants(ctx):
        assert np.array_equal(mi.match_counts(frames, ratio=ratio, algo=_lib.ALGO_TENSOR), exact), (seed, cl)
    mi.close()
    assert np.array_equal(exact[0], [oracle.orb_pair_count_knn(frames[0], m, ratio) for m in map_des])


def test_orb_candidate_overflow_is_repaired_on_the_device(ctx):
    map_des, q = synth.orb_map_and_query(3, n_images=12, nm=400, nq=500)
    frames = [q, synth.orb_flip(np.random.default_rng(1), map_des[5][:300], 20)]
    m = _lib.MapIndex(ctx, "ORB", map_des)
    want = m.match_counts(frames, algo=_lib.ALGO_SIMT)
    assert want.sum() > 300
    repaired0 = ctx.stats()["repaired_pairs"]
    ctx.set_option(3, 16)
    try:
        for cl in tensor_launch_variants(ctx):
            assert np.array_equal(m.match_counts(frames, algo=_lib.ALGO_TENSOR), want), cl
        assert ctx.stats()["repaired_pairs"] > repaired0, "the dirty-pair repair did not run"
    finally:
        ctx.set_option(3, 0)
    m.close()


def test_orb_throughput_shape_tensor_equals_simt(ctx):
    """The batch shape the ORB bench times (frames x images x 500 x 500), reduced: tensor cores vs SIMT, bit-exact."""
    rng = np.random.default_rng(2)
    map_des = [synth.orb_like(rng, 500) for _ in range(60)]
    frames = [synth.orb_like(rng, 500) for _ in range(12)]
    for f in range(12):
        frames[f][:200] = synth.orb_flip(rng, map_des[(7 * f) % 60][:200], 40)
    m = _lib.MapIndex(ctx, "ORB", map_des)
    tc = m.match_counts(frames, algo=_lib.ALGO_TENSOR)
    assert np.array_equal(tc, m.match_counts(frames, algo=_lib.ALGO_SIMT))
    assert all(tc[f, (7 * f) % 60] >= 100 for f in range(12))
    assert ctx.stats()["orb_candidates"] > 0
    m.close()


# ------------------------------------------------------------------------------------------------
# SURF (float L2; tolerance: rows whose ratio is within 1e-6 relative of the threshold may flip)
# ------------------------------------------------------------------------------------------------
def test_surf_vs_oracle_with_tie_tolerance(ctx):
    rng = np.random.default_rng(4)
    map_des = [synth.surf_like(rng, n) for n in (300, 64, 1024, 2, 1, 513)]
    frames = [synth.surf_like(rng, 400), synth.surf_like(rng, 1024)]
    frames[0][:120] = synth.surf_noisy_copy(rng, map_des[0][:120])
    frames[1][:300] = synth.surf_noisy_copy(rng, map_des[2][:300])
    m = _lib.MapIndex(ctx, "SURF", map_des)
    got = m.match_counts(frames)
    m.close()
    for f, q in enumerate(frames):
        for i, d in enumerate(map_des):
            want, ambiguous = oracle.surf_pair_count(q, d, tie_rel=1e-6)
            assert abs(int(got[f, i]) - want) <= ambiguous, (f, i, got[f, i], want, ambiguous)
    assert got[0, 0] >= 100 and got[1, 2] >= 250 and got[:, 4].sum() == 0


# ------------------------------------------------------------------------------------------------
# BOW
# ------------------------------------------------------------------------------------------------
def check_bow_against_oracle(scores, words, frames, locs, W, atol=1e-6):
    """BOWMatch parity (Matcher.py:232-259), unconditional: (a) a word may differ from the exact float64 argmin only where the
    best and second-best squared distances are within 1e-5 relative (float32 near-ties -- scipy's own float32 vq flips those
    too); (b) every score equals the reference's arithmetic (histogram, *idf, L2 normalise, dot) on the words the GPU chose,
    to 1e-6 (BASELINE.md's gate); (c) frames all of whose words equal the exact ones are thereby checked against the exact
    words' scores.  Returns the number of (location, row) near-tie words that differed."""
    allq = np.vstack([f for f in frames if f is not None and len(f)])
    col = 0
    flipped = 0
    for l, (im, idf, voc) in enumerate(locs):
        ow, gap = oracle.bow_words(allq, voc)
        differ = words[l] != ow
        assert np.all(gap[differ] < 1e-5), ("location %d: a word differs from the exact argmin outside the near-tie band" % l,
                                            gap[differ].max())
        flipped += int(differ.sum())
        pos = 0
        for f, q in enumerate(frames):
            n = 0 if q is None else len(q)
            want = oracle.bow_score_from_words(words[l, pos:pos + n], idf, im, W)
            assert np.allclose(scores[f, col:col + len(im)], want[0], rtol=0, atol=atol), (l, f)
            pos += n
        col += len(im)
    return flipped


def test_bow_golden_reference(ctx, golden_dir):
    z = np.load(os.path.join(golden_dir, "bow_case.npz"))
    W = int(z["num_words"])
    # the reference's own histogram width is self.numWords; the golden index was built with numWords=200
    bow = _lib.BowIndex(ctx, [(z["im_features"], z["idf"], z["voc"])], z["im_features"].shape[1])
    scores, words = bow.score([z["qdes"]], want_words=True)
    bow.close()
    assert W == z["im_features"].shape[1]
    check_bow_against_oracle(scores, words, [z["qdes"].astype(np.float32)], [(z["im_features"], z["idf"], z["voc"])], W)
    # against the reference run itself: scipy's words (float32 vq) may differ from ours on near-ties only, and where they
    # are the same the reference's score must be reproduced to 1e-6
    ow, gap = oracle.bow_words(z["qdes"], z["voc"])
    differ = words[0] != z["words"]
    assert np.all(gap[differ] < 1e-5), "visual words may differ from scipy's only on fp32 near-ties"
    ref_on_our_words = oracle.bow_score_from_words(words[0], z["idf"], z["im_features"], W)
    assert np.allclose(scores, ref_on_our_words, rtol=0, atol=1e-6)
    if not differ.any():
        assert np.allclose(scores, z["score"], rtol=0, atol=1e-6)


def test_bow_config5_shape_vs_oracle(ctx):
    """BASELINE config 5's shape per location: numWords = 10 000 (Matcher.py:46), 100 images, 300 query descriptors per frame,
    4 locations (79 word tiles each, the word-range splits of the tensor-core pass, the accept / rescan logic at full width):
    clustered frames (re-observed code-book-like points, what BoW is used on) and a random frame (near-ties everywhere)."""
    W, L, N = 10000, 4, 100
    locs = []
    for l in range(L):
        rng = np.random.default_rng(900 + l)
        rows = W - 37 * l                                    # kmeans drops empty clusters: fewer rows than numWords
        voc = np.clip(synth.sift_like(rng, rows) + rng.normal(0, 0.3, size=(rows, 128)), 0, 255).astype(np.float32)
        im = (rng.random((N, W), dtype=np.float32) * (rng.random((N, W)) < 0.03)).astype(np.float32)
        im /= np.maximum(np.linalg.norm(im, axis=1, keepdims=True), 1e-9)
        locs.append((im.astype(np.float32), (rng.random(W, dtype=np.float32) + 0.5), voc))
    rng = np.random.default_rng(99)
    frames = []
    for f in range(5):
        src = locs[f % L][2]
        frames.append(np.clip(np.rint(src[rng.choice(len(src), 300)] + rng.normal(0, 6.0, size=(300, 128))), 0, 255).astype(np.float32))
    frames.append(synth.sift_like(rng, 300))
    bow = _lib.BowIndex(ctx, locs, W)
    ctx.profile_enable(True)
    scores, words = bow.score(frames, want_words=True)
    _, n_tc = ctx.profile_read("bow_vq_tc")
    ctx.profile_enable(False)
    assert n_tc >= 1, "the tensor-core kernel did not run"
    assert scores.shape == (6, L * N) and words.shape == (L, 6 * 300)
    check_bow_against_oracle(scores, words, frames, locs, W)
    # a clustered frame finds its words back: most rows map to the centroid they were drawn around
    ow, _ = oracle.bow_words(frames[0], locs[0][2])
    assert np.array_equal(words[0, :300], ow) or (words[0, :300] != ow).sum() <= 3
    # one frame at a time gives the same rows as the batch (different work decomposition, same results)
    s1, w1 = bow.score(frames[2:3], want_words=True)
    assert np.array_equal(w1, words[:, 600:900]) and np.allclose(s1, scores[2:3], rtol=0, atol=1e-7)
    # and the float32 SIMT kernel agrees except on near-ties
    ctx.set_option(2, 1)
    try:
        s_simt, w_simt = bow.score(frames, want_words=True)
    finally:
        ctx.set_option(2, 0)
    allq = np.vstack(frames)
    for l in range(L):
        _, gap = oracle.bow_words(allq, locs[l][2])
        assert np.all(gap[words[l] != w_simt[l]] < 1e-5)
    bow.close()


def test_bow_multi_location_vs_oracle(ctx):
    W = 256
    locs, dess = [], []
    for l in range(3):
        im, idf, voc, des = synth.bow_index(50 + l, n_images=6 + l, num_words=W, voc_rows=W - 3 * l, des_per_image=120)
        locs.append((im, idf, voc))
        dess.append(des)
    rng = np.random.default_rng(8)
    frames = [synth.sift_noisy_copy(rng, dess[1][2], 3.0), synth.sift_like(rng, 77), None]
    bow = _lib.BowIndex(ctx, locs, W)
    scores, words = bow.score(frames, want_words=True)
    bow.close()
    check_bow_against_oracle(scores, words, frames, locs, W)
    # frame 0 re-observes image 2 of location 1: it must win there
    off1 = len(locs[0][0])
    assert int(np.argmax(scores[0, off1:off1 + len(locs[1][0])])) == 2


def test_bow_tensor_core_path_equals_simt_and_oracle(ctx):
    """Word assignment on tensor cores (two u8 planes of the 16-bit fixed-point codebook + float32 resolve) gives the
    same words as the float32 SIMT kernel and as the oracle (up to float32 near-ties); non-integer query rows and
    out-of-range centroids take the SIMT kernel on the device."""
    W = 640
    rng = np.random.default_rng(21)
    locs = []
    for rows in (640, 500, 129):
        voc = np.clip(synth.sift_like(rng, rows) + rng.normal(0, 0.4, size=(rows, 128)), 0, 255).astype(np.float32)
        voc[5] = voc[4]                      # an exact tie between two words: lowest index must win
        im = rng.random((7, W), dtype=np.float32)
        im /= np.linalg.norm(im, axis=1, keepdims=True)
        locs.append((im, (rng.random(W, dtype=np.float32) + 0.5), voc))
    clustered = np.clip(np.rint(locs[0][2][rng.choice(640, 300)] + rng.normal(0, 5.0, size=(300, 128))), 0, 255).astype(np.float32)
    clustered[:3] = np.rint(locs[0][2][4])   # rows sitting on the tied pair
    frames = [clustered, synth.sift_like(rng, 50), synth.sift_like(rng, 1), None]
    bow = _lib.BowIndex(ctx, locs, W)
    ctx.profile_enable(True)
    scores, words = bow.score(frames, want_words=True)
    _, n_tc = ctx.profile_read("bow_vq_tc")
    ctx.profile_enable(False)
    assert n_tc == 1, "the tensor-core kernel did not run"
    ctx.set_option(2, 1)
    scores_simt, words_simt = bow.score(frames, want_words=True)
    ctx.set_option(2, 0)
    allq = np.vstack([f for f in frames if f is not None])
    for l, (_, _, voc) in enumerate(locs):
        ow, gap = oracle.bow_words(allq, voc)
        for got in (words[l], words_simt[l]):
            differ = got != ow
            assert np.all(gap[differ] < 1e-5)
        same = words[l] == words_simt[l]
        assert np.all(gap[~same] < 1e-5)
    assert np.array_equal(words[0][:3], [4, 4, 4])
    if np.array_equal(words, words_simt):
        assert np.allclose(scores, scores_simt, rtol=0, atol=1e-7)
    # non-integer query rows: rejected by the u8 ingest on the device, the SIMT kernel assigns the words
    frac = [f + 0.25 if f is not None else None for f in frames]
    s_frac, w_frac = bow.score(frac, want_words=True)
    ctx.set_option(2, 1)
    s_ref, w_ref = bow.score(frac, want_words=True)
    ctx.set_option(2, 0)
    assert np.array_equal(w_frac, w_ref) and np.array_equal(s_frac, s_ref)
    bow.close()
    # a centroid outside [0, 256): no planes are built for this index, SIMT only
    bad_locs = [(locs[0][0], locs[0][1], locs[0][2] - 3.0)]
    bow2 = _lib.BowIndex(ctx, bad_locs, W)
    ctx.profile_enable(True)
    s2, w2 = bow2.score(frames[:2], want_words=True)
    _, n_tc = ctx.profile_read("bow_vq_tc")
    ctx.profile_enable(False)
    assert n_tc == 0
    ow, gap = oracle.bow_words(np.vstack(frames[:2]), bad_locs[0][2])
    assert np.all(gap[w2[0] != ow] < 1e-5)
    bow2.close()


def test_kmeans_matches_scipy_lloyd(ctx):
    """mcl_kmeans = scipy.cluster.vq._kmeans(obs, guess, thresh): same iterations from the same initial code book.
    Observations are integer valued, so member sums are exact and the centroids agree to float32 rounding unless an
    assignment sat on a float32 near-tie (none on this data)."""
    from scipy.cluster.vq import kmeans
    rng = np.random.default_rng(77)
    centers = synth.sift_like(rng, 60)
    obs = np.clip(np.rint(centers[rng.integers(0, 60, size=5000)] + rng.normal(0, 9.0, size=(5000, 128))), 0, 255).astype(np.float32)
    guess = obs[rng.choice(5000, size=48, replace=False)].copy()
    guess[7] = guess[3]     # duplicate initial centroid -> an empty cluster that must be dropped (lowest index wins ties)
    ref_book, ref_dist = kmeans(obs, guess, thresh=1e-5)
    book, dist, iters = ctx.kmeans(obs, guess, thresh=1e-5)
    assert book.shape == ref_book.shape and book.shape[0] < 48
    assert iters >= 2
    assert np.allclose(book, ref_book, rtol=0, atol=1e-3)
    assert abs(dist - float(ref_dist)) <= 1e-4 * float(ref_dist)
    book8, dist8, _ = ctx.kmeans(obs.astype(np.uint8), guess, thresh=1e-5)
    assert np.array_equal(book8, book) and dist8 == dist


# ------------------------------------------------------------------------------------------------
# Laplacian variance
# ------------------------------------------------------------------------------------------------
def test_laplacian_sums_exact_and_var(ctx):
    """Sums are exact integers; the variance is numpy's two-pass pairwise variance, bit for bit = what the reference's
    cv2.Laplacian(img, CV_64F).var() returns (analyze.py:441-445), including on blurred images (variance below 200: the
    value that reaches out.txt through the blur weight)."""
    import cv2
    rng = np.random.default_rng(2)
    imgs = [rng.integers(0, 256, size=s, dtype=np.uint8) for s in [(600, 800), (1, 1), (1, 7), (9, 1), (2, 2), (33, 129),
                                                                    (240, 320), (481, 641), (3, 5), (130, 1), (600, 801)]]
    for k in range(1, 9):
        imgs.append(cv2.GaussianBlur(imgs[0 if k % 2 else 6], (0, 0), 0.5 * k + 0.3))
    imgs.append(np.full((50, 60), 200, np.uint8))
    sums = ctx.laplacian_sums(imgs)
    var = ctx.laplacian_var(imgs)
    low = 0
    for im, s, v in zip(imgs, sums, var):
        s1, s2, n = oracle.laplacian_sums(im)
        assert (int(s[0]), int(s[1])) == (s1, s2)
        ref = oracle.cv2_laplacian_var(im)
        assert v == ref, (im.shape, v, ref)
        assert v == oracle.laplacian_var(im)
        low += int(ref < 200)
    assert low >= 5
    # non-contiguous rows (stride > width), one image at a time and in a mixed-size batch
    big = rng.integers(0, 256, size=(100, 300), dtype=np.uint8)
    view = big[:, 17:217]
    assert ctx.laplacian_var([view])[0] == oracle.cv2_laplacian_var(np.ascontiguousarray(view))
    again = ctx.laplacian_var(imgs[::-1])
    assert np.array_equal(again, var[::-1])


# ------------------------------------------------------------------------------------------------
# particle filter update
# ------------------------------------------------------------------------------------------------
def _state(rng, L, I):
    return [[float(rng.integers(1, 400)), list(rng.random(I))] for _ in range(L)]


@pytest.mark.parametrize("L,I", [(7, 25), (3, 25), (64, 50)])
def test_filter_step_bit_exact(ctx, L, I):
    import copy
    rng = np.random.default_rng(L * 100 + I)
    for trial in range(12):
        prev = _state(rng, L, I)
        cur = _state(rng, L, I)
        cmd = "lrfbs"[trial % 5]
        blur = [0.0, 37.5, 199.999, 200.0, 200.1, 913.2][trial % 6]
        if cmd == "f" and I == 25:  # make both forward branches fire
            bc = 0 if trial % 2 else L - 1
            prev[bc][0] = 1000.0
            prev[bc][1][3 if trial % 2 else 20] = 2.0
        o_prev = copy.deepcopy(prev)
        if I == 25:
            adj, best = oracle.filter_step(cmd, o_prev, cur, blur)
        else:  # oracle is written for 15-degree steps too; generalised sizes share the code
            adj, best = oracle.filter_step(cmd, o_prev, cur, blur)
        p = np.array([[c[0]] + c[1] for c in prev])
        c = np.array([[x[0]] + x[1] for x in cur])
        prev2, out, gbest = ctx.filter_step(_lib.F_ALL, cmd, blur, p, c)
        want = np.array([[a[0]] + a[1] for a in adj])
        wprev = np.array([[a[0]] + a[1] for a in o_prev])
        assert np.array_equal(out, want), (cmd, blur)
        assert np.array_equal(prev2, wprev)
        assert gbest == tuple(best)


def test_filter_forward_factor_table(ctx):
    """0.05*|sin(k*15*180/pi)| (sic, analyze.py:331) through the kernel: bump of the neighbour's total."""
    for k in range(1, 12):
        prev = [[1.0, [0.0] * 25] for _ in range(3)]
        prev[0][0] = 5.0
        prev[0][1][k] = 1.0
        p = np.array([[c[0]] + c[1] for c in prev])
        prev2, _, _ = ctx.filter_step(_lib.F_COMMAND, "f", 0.0, p, p)
        assert prev2[1, 0] == 1.0 * (1 + oracle.forward_factor(k))


def _typed_state(rng, L, I, kinds):
    """Rows whose scalars are Python floats, np.float64 or np.float32 (what the colour method feeds the filter)."""
    rows = []
    for l in range(L):
        k = kinds[l % len(kinds)]
        conv = {"py": float, "f64": np.float64, "f32": np.float32}[k]
        rows.append([conv(rng.integers(1, 400) + rng.random()), [conv(v) for v in rng.random(I)]])
    return rows


def _tags(rows):
    def t(x):
        return _lib.T_F32 if isinstance(x, np.float32) else (_lib.T_F64 if isinstance(x, np.float64) else _lib.T_PY)
    return np.array([[t(r[0]), t(r[1][0])] for r in rows], dtype=np.int8)


@pytest.mark.parametrize("kinds_prev,kinds_cur", [(("py",), ("f32",)), (("f32",), ("f32", "py")), (("f64",), ("f32", "py")),
                                                  (("f32", "py", "f64"), ("py", "f32", "f32", "f64"))])
def test_filter_step_typed_follows_numpy_scalar_promotion(ctx, kinds_prev, kinds_cur):
    """The reference's filter is plain Python operators over lists; with np.float32 / np.float64 elements (colour method)
    every multiply and add follows numpy's promotion rules.  The oracle is those operators; the kernel must give the same
    values AND the same result types, bit for bit."""
    import copy
    L, I = 7, 25
    rng = np.random.default_rng(len(kinds_prev) * 10 + len(kinds_cur))
    for trial in range(20):
        prev = _typed_state(rng, L, I, kinds_prev)
        cur = _typed_state(rng, L, I, kinds_cur)
        cmd = "lrfbs"[trial % 5]
        blur = [np.float64(0.0), np.float64(37.5), np.float64(199.999), np.float64(200.1), 913.2, 12.25][trial % 6]
        if cmd == "f":
            bc = 0 if trial % 2 else L - 1
            prev[bc][0] = type(prev[bc][0])(1000.0)
            prev[bc][1][3 if trial % 2 else 20] = type(prev[bc][1][0])(2.0)
        o_prev = copy.deepcopy(prev)
        adj, best = oracle.filter_step(cmd, o_prev, cur, blur)
        p = np.array([[c[0]] + c[1] for c in prev], dtype=np.float64)
        c = np.array([[x[0]] + x[1] for x in cur], dtype=np.float64)
        bt = _lib.T_F64 if isinstance(blur, np.float64) else _lib.T_PY
        prev2, out, ot, gbest = ctx.filter_step_typed(_lib.F_ALL, cmd, float(blur), bt, p, _tags(prev), c, _tags(cur))
        want = np.array([[a[0]] + a[1] for a in adj], dtype=np.float64)
        assert np.array_equal(out, want), (cmd, blur, trial)
        assert np.array_equal(ot, _tags(adj)), (cmd, blur, trial)
        assert np.array_equal(prev2, np.array([[a[0]] + a[1] for a in o_prev], dtype=np.float64))
        assert gbest == tuple(best)


# ------------------------------------------------------------------------------------------------
# colour histograms: chi-squared search
# ------------------------------------------------------------------------------------------------
def test_chi2_equals_reference_searcher_bitwise(ctx, golden_dir):
    g = np.load(os.path.join(golden_dir, "color_case.npz"))
    got = ctx.chi2(g["q"], g["m"])
    assert got.dtype == np.float32
    assert np.array_equal(got.view(np.uint32), g["chi"].view(np.uint32))


@pytest.mark.parametrize("bins", [1, 5, 8, 9, 64, 127, 128, 129, 200, 511, 512, 513, 1000, 4096])
def test_chi2_any_bin_count_follows_numpy_pairwise_order(ctx, bins):
    rng = np.random.default_rng(bins)
    q = (rng.random((5, bins)) * 255).astype(np.float32)
    m = (rng.random((9, bins)) * 255).astype(np.float32)
    m[0] = q[0]                      # identical histograms: distance 0
    q[1] = 0
    m[1] = 0                         # empty bins on both sides: 0 / eps = 0, not NaN
    m[2, ::3] = 0
    got = ctx.chi2(q, m)
    want = oracle.chi2_matrix(q, m)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert got[0, 0] == 0 and got[1, 1] == 0


def test_chi2_large_batch_and_empty(ctx):
    rng = np.random.default_rng(8)
    hist = lambda n: np.floor(rng.gamma(0.5, 40.0, size=(n, 512))).clip(0, 255).astype(np.float32)
    q, m = hist(40), hist(700)
    got = ctx.chi2(q, m)
    want = oracle.chi2_matrix(q, m)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert ctx.chi2(q[:0], m).shape == (0, 700) and ctx.chi2(q, m[:0]).shape == (40, 0)
    with pytest.raises(_lib.MclError):
        ctx.chi2(q, m[:, :100])
