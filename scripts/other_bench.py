"""Side benchmarks of the other BASELINE.json configs (ORB cfg1, SIFT cfg2, SURF cfg4, BOW cfg5, Laplacian), each with its
roofline and the cv2/scipy CPU routine timed beside it.  python scripts/other_bench.py [names...]  -> JSON lines."""
import json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "oracle"))
import numpy as np
import pymcl_b200
from pymcl_b200 import _lib, synth
import oracle

ctx = _lib.Context(0)
names = sys.argv[1:] or ["orb", "orb_batch", "sift_cfg2", "surf", "bow", "bow_batch", "kmeans", "laplacian", "chi2"]
POPC_PEAK = None


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.sync()
    return (time.perf_counter() - t0) / reps


def prof(fn, slot, reps=10, warm=3):
    for _ in range(warm):
        fn()
    for k in _lib.PROFILE_SLOTS:      # drop whatever earlier measurements left in the other slots
        ctx.profile_read(k)
    ctx.profile_enable(True)
    for _ in range(reps):
        fn()
    ms, n = ctx.profile_read(slot)
    ctx.profile_enable(False)
    return ms / reps


def cpu_rate(fn, budget=5.0):
    fn()
    t0 = time.perf_counter(); n = 0
    while time.perf_counter() - t0 < budget:
        fn(); n += 1
    return n / (time.perf_counter() - t0)


out = []
if "orb" in names or "orb_batch" in names:
    POPC_PEAK, _ = ctx.microbench(1, 20000)
if "orb" in names:   # cfg1: one (500,32) query vs 50 x (500,32)
    import cv2
    map_des, q = synth.orb_map_and_query(1)
    m = _lib.MapIndex(ctx, "ORB", map_des)
    for mode, tag in ((_lib.KNN2_RATIO, "knn"), (_lib.CROSSCHECK, "crosscheck")):
        e2e = timeit(lambda: m.match_counts([q], mode=mode), reps=50)
        k_ms = prof(lambda: m.match_counts([q], mode=mode), "orb", reps=50)
        popc = 500 * 500 * 8 * 50 * (1 if mode == _lib.KNN2_RATIO else 2)
        f = (lambda: [oracle.cv2_orb_pair_count_knn(q, d) for d in map_des]) if mode == _lib.KNN2_RATIO else \
            (lambda: [oracle.cv2_orb_pair_count_crosscheck(q, d) for d in map_des])
        cpu = cpu_rate(f, 3.0) * 50
        out.append({"config": "cfg1 ORB 1x50 " + tag, "e2e_pairs_per_s": 50 / e2e, "e2e_ms": e2e * 1e3, "kernel_ms": k_ms,
                    "popc_per_s": popc / (k_ms * 1e-3), "frac_popc_peak": popc / (k_ms * 1e-3) / POPC_PEAK,
                    "cpu_pairs_per_s": cpu, "cpu_threads": cv2.getNumThreads()})
    m.close()
if "orb_batch" in names:   # throughput shape: 200 frames x 400 images of 500 rows
    rng = np.random.default_rng(1)
    map_des = [synth.orb_like(rng, 500) for _ in range(400)]
    frames = [synth.orb_like(rng, 500) for _ in range(200)]
    m = _lib.MapIndex(ctx, "ORB", map_des)
    for mode, tag in ((_lib.KNN2_RATIO, "knn"), (_lib.CROSSCHECK, "crosscheck")):
        k_ms = prof(lambda: m.match_counts(frames, mode=mode), "orb", reps=3, warm=1)
        popc = 500 * 500 * 8 * 400 * 200 * (1 if mode == _lib.KNN2_RATIO else 2)
        out.append({"config": "ORB 200x400 " + tag, "kernel_ms": k_ms, "pairs_per_s": 80000 / (k_ms * 1e-3),
                    "popc_per_s": popc / (k_ms * 1e-3), "frac_popc_peak": popc / (k_ms * 1e-3) / POPC_PEAK,
                    "popc_peak_measured": POPC_PEAK})
    m.close()
if "sift_cfg2" in names:   # cfg2: 200 frames x 256 rows vs 8 loc x 50 img x 256 rows
    map_des, frames = synth.sift_map_and_frames(2, 400, 256, 200, 256)
    m = _lib.MapIndex(ctx, "SIFT", map_des)
    res = {}
    for algo, tag in ((_lib.ALGO_TENSOR, "tc2"), (_lib.ALGO_SIMT, "simt")):
        e2e = timeit(lambda: m.match_counts(frames, algo=algo), reps=5, warm=2)
        res[tag + "_e2e_ms"] = e2e * 1e3
    k_ms = prof(lambda: m.match_counts(frames, algo=_lib.ALGO_TENSOR), "sift_tc", reps=5, warm=2)
    ops = 2.0 * 256 * 256 * 128 * 80000
    import cv2
    cpu = cpu_rate(lambda: [oracle.cv2_l2_pair_count(frames[0], d) for d in map_des[:50]], 3.0) * 50
    res.update({"config": "cfg2 SIFT 200x400 (256x256)", "kernel_ms": k_ms, "kernel_TOPs": ops / (k_ms * 1e-3) / 1e12,
                "e2e_pairs_per_s": 80000 / (res["tc2_e2e_ms"] * 1e-3), "cpu_pairs_per_s": cpu, "cpu_threads": cv2.getNumThreads()})
    out.append(res)
    m.close()
if "surf" in names:   # cfg4 shape: 1024 x 1024 x 64 f32, one frame vs 30 images (a DOR step) and vs 175
    rng = np.random.default_rng(4)
    map_des = [synth.surf_like(rng, 1024) for _ in range(175)]
    q = synth.surf_like(rng, 1024)
    m = _lib.MapIndex(ctx, "SURF", map_des)
    mask = np.zeros((1, 175), np.uint8); mask[0, :30] = 1
    pair_peak, _ = ctx.microbench(2, 20000)     # (FADD, FFMA) element pairs per second: the direct-difference form's bound
    frames40 = [synth.surf_like(rng, 1024) for _ in range(40)]
    for tag, kw, n, qq in (("175 images", {}, 175, [q]), ("30 of 175 (DOR mask)", {"mask": mask}, 30, [q]),
                           ("40 frames x 175 images (throughput shape)", {}, 175 * 40, frames40)):
        e2e = timeit(lambda: m.match_counts(qq, **kw), reps=20 if len(qq) == 1 else 3)
        k_ms = prof(lambda: m.match_counts(qq, **kw), "surf", reps=20 if len(qq) == 1 else 3)
        elems = 1024.0 * 1024 * 64 * n
        out.append({"config": "cfg4 SURF 1024x1024x64: " + tag, "e2e_ms": e2e * 1e3, "kernel_ms": k_ms,
                    "tflops_at_3_flop_per_element": 3.0 * elems / (k_ms * 1e-3) / 1e12,
                    "element_pairs_per_s": elems / (k_ms * 1e-3), "fp32_pair_peak_measured": pair_peak,
                    "frac_of_fp32_peak": elems / (k_ms * 1e-3) / pair_peak, "pairs_per_s_e2e": n / e2e})
    import cv2
    cpu = cpu_rate(lambda: [oracle.cv2_l2_pair_count(q, d) for d in map_des[:10]], 3.0) * 10
    out[-1]["cpu_pairs_per_s"] = cpu
    m.close()
if "bow" in names:   # cfg5 scaled: 16 locations x 100 images, W = 10000, Nq = 300
    L = 16
    locs = []
    for l in range(L):
        rng = np.random.default_rng(500 + l)
        voc = np.clip(synth.sift_like(rng, 10000) + rng.normal(0, 0.3, size=(10000, 128)), 0, 255).astype(np.float32)
        im = rng.random((100, 10000), dtype=np.float32) * (rng.random((100, 10000)) < 0.03)
        im /= np.maximum(np.linalg.norm(im, axis=1, keepdims=True), 1e-9)
        idf = rng.random(10000, dtype=np.float32) + 0.5
        locs.append((im.astype(np.float32), idf, voc))
    bow = _lib.BowIndex(ctx, locs, 10000)
    rng = np.random.default_rng(9)
    slots = ("bow_vq_tc", "bow_resolve", "bow_vq", "bow_score")

    def run(q, reps=5, warm=2):
        for _ in range(warm):
            bow.score([q])
        t0 = time.perf_counter()
        for _ in range(reps):
            bow.score([q])
        ctx.sync()
        dt = (time.perf_counter() - t0) / reps      # end to end, profiling off
        for k in slots:
            ctx.profile_read(k)
        ctx.profile_enable(True)
        for _ in range(reps):
            bow.score([q])
        r = {k: ctx.profile_read(k)[0] / reps for k in slots}
        ctx.profile_enable(False)
        r["e2e_ms"] = dt * 1e3
        return r

    # (a) clustered queries: re-observed codebook-like points (what BoW is used on); (b) random queries (adversarial:
    # best and second-best distances nearly tie, so most rows take the exact float32 route)
    q_clu = np.clip(np.rint(locs[3][2][rng.choice(10000, 300)] + rng.normal(0, 6.0, size=(300, 128))), 0, 255).astype(np.float32)
    q_rnd = synth.sift_like(rng, 300)
    from scipy.cluster.vq import vq
    cpu = cpu_rate(lambda: oracle.scipy_bow_score(q_clu, locs[0][2], locs[0][1], locs[0][0], 10000), 3.0)
    for tag, q in (("clustered", q_clu), ("random", q_rnd)):
        r = run(q)
        ctx.set_option(2, 1)
        r_simt = run(q)
        ctx.set_option(2, 0)
        out.append({"config": "cfg5 BOW 1 frame x %d locations x 100 images (W=10000, Nq=300), %s queries" % (L, tag),
                    "e2e_ms": r["e2e_ms"], "vq_tc_ms": r["bow_vq_tc"], "vq_resolve_ms": r["bow_resolve"],
                    "vq_simt_fallback_ms": r["bow_vq"], "score_ms": r["bow_score"],
                    "vq_tc_fp16_TFLOPs_algorithmic": 2.0 * 300 * 10000 * 128 * L / (max(r["bow_vq_tc"], 1e-9) * 1e-3) / 1e12,
                    "locations_per_s_e2e": L / (r["e2e_ms"] * 1e-3),
                    "simt_only_e2e_ms": r_simt["e2e_ms"], "simt_only_vq_ms": r_simt["bow_vq"],
                    "cpu_locations_per_s": cpu})
    bow.close()
if "kmeans" in names:   # createIndex's code book training: 25 images x 2000 descriptors, k = 10000 (Matcher.py:46,124)
    from scipy.cluster.vq import kmeans as sp_kmeans
    rng = np.random.default_rng(8)
    centers = synth.sift_like(rng, 12000)
    obs = np.clip(np.rint(centers[rng.integers(0, 12000, size=50000)] + rng.normal(0, 7.0, size=(50000, 128))), 0, 255).astype(np.float32)
    guess = obs[rng.choice(50000, size=10000, replace=False)]
    ctx.kmeans(obs[:2000], guess[:100])
    t0 = time.perf_counter(); book, dist, iters = ctx.kmeans(obs, guess); t1 = time.perf_counter()
    sub = 4
    t2 = time.perf_counter(); rb, rd = sp_kmeans(obs[::sub], guess[::sub], thresh=1e-5); t3 = time.perf_counter()
    out.append({"config": "k-means code book, 50000 x 128 observations, k = 10000 (createIndex)", "gpu_s": t1 - t0, "iterations": iters,
                "words_kept": int(book.shape[0]), "distortion": dist,
                "scipy_s_on_1/%d of obs and 1/%d of k" % (sub, sub): t3 - t2,
                "scipy_extrapolated_s": (t3 - t2) * sub * sub})
if "laplacian" in names:
    import cv2
    rng = np.random.default_rng(6)
    imgs = [rng.integers(0, 256, size=(600, 800), dtype=np.uint8) for _ in range(200)]
    e2e = timeit(lambda: ctx.laplacian_var(imgs), reps=5, warm=2)
    k_ms = prof(lambda: ctx.laplacian_var(imgs), "laplacian", reps=5, warm=2)
    cpu = cpu_rate(lambda: [oracle.cv2_laplacian_var(i) for i in imgs[:20]], 3.0) * 20
    out.append({"config": "Laplacian 200 x 800x600", "e2e_ms": e2e * 1e3, "kernel_ms": k_ms,
                "kernel_GBps": 200 * 480000 / (k_ms * 1e-3) / 1e9, "images_per_s_e2e": 200 / e2e, "cpu_images_per_s": cpu})
if "bow_batch" in names:   # createRawP's shape: many frames against every location in one call
    L, F = 16, 50
    locs = []
    for l in range(L):
        rng = np.random.default_rng(500 + l)
        voc = np.clip(synth.sift_like(rng, 10000) + rng.normal(0, 0.3, size=(10000, 128)), 0, 255).astype(np.float32)
        im = rng.random((100, 10000), dtype=np.float32) * (rng.random((100, 10000)) < 0.03)
        im /= np.maximum(np.linalg.norm(im, axis=1, keepdims=True), 1e-9)
        locs.append((im.astype(np.float32), rng.random(10000, dtype=np.float32) + 0.5, voc))
    bow = _lib.BowIndex(ctx, locs, 10000)
    rng = np.random.default_rng(9)
    frames = [np.clip(np.rint(locs[f % L][2][rng.choice(10000, 300)] + rng.normal(0, 6.0, size=(300, 128))), 0, 255).astype(np.float32)
              for f in range(F)]
    e2e = timeit(lambda: bow.score(frames), reps=3, warm=2)
    r = {k: prof(lambda: bow.score(frames), k, reps=3, warm=1) for k in ("bow_vq_tc", "bow_resolve", "bow_rescan", "bow_vq", "bow_score")}
    out.append({"config": "cfg5 BOW batched: %d frames x %d locations x 100 images (W=10000, Nq=300)" % (F, L), "e2e_ms": e2e * 1e3,
                "vq_tc_ms": r["bow_vq_tc"], "vq_resolve_ms": r["bow_resolve"], "vq_rescan_ms": r["bow_rescan"], "vq_simt_fallback_ms": r["bow_vq"], "score_ms": r["bow_score"],
                "vq_tc_fp16_TFLOPs_algorithmic": 2.0 * F * 300 * 10000 * 128 * L / (r["bow_vq_tc"] * 1e-3) / 1e12,
                "frame_locations_per_s_e2e": F * L / e2e})
    bow.close()
if "chi2" in names:   # colour method: 200 frame histograms vs 7 x 25 map histograms (512 bins)
    rng = np.random.default_rng(12)
    hist = lambda n: np.floor(rng.gamma(0.5, 40.0, size=(n, 512))).clip(0, 255).astype(np.float32)
    q, m = hist(200), hist(175)
    e2e = timeit(lambda: ctx.chi2(q, m), reps=20)
    k_ms = prof(lambda: ctx.chi2(q, m), "chi2", reps=20)
    cpu = cpu_rate(lambda: oracle.chi2_matrix(q[:4], m), 3.0) * 4 * 175
    out.append({"config": "Color chi-squared 200 x 175 histograms (512 bins)", "e2e_ms": e2e * 1e3, "kernel_ms": k_ms,
                "pairs_per_s_e2e": 200 * 175 / e2e, "kernel_GBps_from_L2": 200 * 175 * 2 * 2048 / (k_ms * 1e-3) / 1e9,
                "cpu_pairs_per_s_numpy_1_thread": cpu})
for o in out:
    print(json.dumps(o))
