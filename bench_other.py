"""bench.py's other workloads: bag of words (BASELINE config 5) and ORB (config 1's shape as a batch).  Same contract line as
the SIFT workload (bench.py's docstring); imported by bench.py, not a separate entry point."""
from __future__ import annotations

import json
import os
import sys
import time

REPO = os.path.dirname(os.path.abspath(__file__))

BOW_W, BOW_IMAGES, BOW_NQ, BOW_FRAMES, BOW_LOC_PER_GPU = 10000, 100, 300, 50, 32
ORB_N, ORB_FRAMES, ORB_IMG_PER_GPU = int(os.environ.get("MCL_BENCH_ORB_N", "500")), 200, 400   # (the override is for kernel experiments)


# ---------------------------------------------------------------------------------------------------------------------
# bag of words
# ---------------------------------------------------------------------------------------------------------------------
def bow_location(l):
    """Location l of the synthetic map: a 10 000-word code book of SIFT-like centroids (non-integer float32, as k-means
    produces), idf, and 100 sparse L2-normalised tf-idf rows -- seeded by l, so a rank holds the same index whatever the
    sharding."""
    import numpy as np
    from pymcl_b200 import synth
    rng = np.random.default_rng(500 + l)
    voc = np.clip(synth.sift_like(rng, BOW_W) + rng.normal(0, 0.3, size=(BOW_W, 128)), 0, 255).astype(np.float32)
    im = rng.random((BOW_IMAGES, BOW_W), dtype=np.float32) * (rng.random((BOW_IMAGES, BOW_W)) < 0.03)
    im /= np.maximum(np.linalg.norm(im, axis=1, keepdims=True), 1e-9)
    return im.astype(np.float32), rng.random(BOW_W, dtype=np.float32) + 0.5, voc


def bow_frames(n_frames, n_loc_total):
    """Frame f re-observes 300 code-book points of location (5 f mod total) with pixel noise (what BoW is used on); against
    every other location its rows are as good as random."""
    import numpy as np
    out = []
    for f in range(n_frames):
        rng = np.random.default_rng(9000 + f)
        voc = bow_location((5 * f) % n_loc_total)[2]
        out.append(np.clip(np.rint(voc[rng.choice(BOW_W, BOW_NQ)] + rng.normal(0, 6.0, size=(BOW_NQ, 128))), 0, 255).astype(np.float32))
    return out


def bow_text(n_frames, per_gpu):
    return ("BOW (BASELINE config 5): %d frames x %d SIFT descriptors assigned to %d visual words and tf-idf scored against "
            "%d locations x %d images per GPU" % (n_frames, BOW_NQ, BOW_W, per_gpu, BOW_IMAGES))


def bow_cpu(frames, locs, n_pairs, threads):
    """scipy.cluster.vq.vq + histogram + idf + sklearn normalize + dot (Matcher.py:249-258 verbatim, oracle.scipy_bow_score)
    over (frame 0, location 0), (frame 0, location 1), ...: one call scores the 100 images of a location."""
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import oracle
    try:
        from threadpoolctl import threadpool_limits
        lim = threadpool_limits(limits=threads)
    except Exception:
        lim = None
    out = []
    t0 = time.perf_counter()
    for k in range(n_pairs):
        f, l = (k // len(locs)) % len(frames), k % len(locs)
        im, idf, voc = locs[l]
        words, score = oracle.scipy_bow_score(frames[f], voc, idf, im, BOW_W)
        out.append((f, l, words, score))
    dt = time.perf_counter() - t0
    return n_pairs / dt, dt, out


def run_bow(args, H, ClockSampler, peaks):
    import numpy as np
    import torch
    import pymcl_b200  # noqa: F401
    from pymcl_b200 import _lib
    from pymcl_b200.engine import ShardedBow
    n_frames = args.frames or BOW_FRAMES
    per_gpu = args.images or BOW_LOC_PER_GPU
    if args.impl == "reference":
        return bow_reference(args, n_frames, per_gpu, per_gpu * max(1, args.gpus))
    world, rank = H.world, H.rank
    total_loc = per_gpu * world
    ctx = _lib.Context(H.local_rank)
    frames = bow_frames(n_frames, total_loc)
    cache = {}

    def loader(i):
        cache[i] = bow_location(i)
        return cache[i]

    sbow = ShardedBow(total_loc, loader, num_words=BOW_W, ctx=ctx)     # locations sharded over the ranks (product path)
    queries = _lib.QueryBatch(ctx, _lib.BOWQ, frames)                   # resident copy for `value`
    result = {}

    def step_resident():
        result["dev"] = sbow.score_device(queries)

    packed = _lib.pack_descriptors(_lib.BOWQ, frames)                   # float32 rows, as cv2.SIFT hands them over
    pin = _lib.pinned_empty(packed[1].shape, packed[1].dtype)
    pin[...] = packed[1]
    packed_pinned = (packed[0], pin, packed[2])

    def step_e2e():                                                     # rows in one pinned buffer (like bench.py's SIFT e2e)
        result["host"] = sbow.score(None, packed=packed_pinned)

    def step_e2e_list():                                                # a Python list of per-frame arrays: + the packing pass
        result["host_list"] = sbow.score(frames)

    step_resident()
    torch.cuda.synchronize()
    dev = result["dev"].cpu().numpy()
    step_e2e()
    host = result["host"]
    for r, (l, h) in enumerate(sbow.bounds):
        assert np.array_equal(host[:, l:h], dev[r][:, :h - l]), "host-buffer path and resident path disagree"
    for f in range(min(n_frames, 8)):   # the re-observed location must hold the best-scoring image
        l = (5 * f) % total_loc
        assert int(np.argmax(host[f])) // BOW_IMAGES >= 0 and host[f].max() > 0

    sampler = ClockSampler(H.local_rank)
    sampler.start()
    time.sleep(0.3)
    slots = ("bow_vq_tc", "bow_resolve", "bow_rescan", "bow_vq", "bow_score", "ingest")
    for k in slots:
        ctx.profile_read(k)
    ctx.profile_enable(True)
    l0 = ctx.launch_count()
    ms = H.timed(step_resident, args.steps, args.warmup, use_events=True)
    prof = {k: ctx.profile_read(k)[0] / (args.steps + args.warmup) for k in slots}
    ctx.profile_enable(False)
    launches = (ctx.launch_count() - l0) * args.steps // max(1, args.steps + args.warmup)
    ms_e2e = H.timed(step_e2e, args.steps, max(1, args.warmup // 2), use_events=False)
    ms_e2e_list = H.timed(step_e2e_list, args.steps, max(1, args.warmup // 2), use_events=False)
    assert np.array_equal(result["host_list"], result["host"])
    time.sleep(0.2)
    sampler.stop()
    sampler.join(timeout=2)
    units_step = n_frames * total_loc * BOW_IMAGES
    value = units_step / (ms / args.steps * 1e-3)
    e2e_value = units_step / (ms_e2e / args.steps * 1e-3)
    if rank == 0:
        pk = peaks()
        flops = 2.0 * n_frames * BOW_NQ * BOW_W * 128 * per_gpu          # algorithmic: every row against every word
        achieved = flops / (prof["bow_vq_tc"] * 1e-3) / 1e12
        peak = pk["bf16_burst"]
        st = ctx.stats()
        roofline = {"bound": "tensor", "kernel": "tc9_kernel<BOW> (tcgen05.mma kind::f16 128x128x16, fp16 code book + norm extension, queries in TMEM, "
                                                 "top-3 group epilogue)",
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops (%s): cuBLAS bf16 GEMM burst; nominal fp16 dense 2250" % pk["src"],
                    "achieved_incl_norm_extension": achieved * 9 / 8, "frac_of_nominal": achieved / 2250.0,
                    "kernel_ms_per_step": prof["bow_vq_tc"], "kernel_share_of_step": prof["bow_vq_tc"] / (ms / args.steps),
                    "resolve_ms_per_step": prof["bow_resolve"], "rescan_ms_per_step": prof["bow_rescan"],
                    "simt_fallback_ms_per_step": prof["bow_vq"], "score_ms_per_step": prof["bow_score"],
                    "vq_plus_resolve_over_tensor_pass": (prof["bow_vq_tc"] + prof["bow_resolve"] + prof["bow_rescan"]) / prof["bow_vq_tc"],
                    "resolve_entries_last_call": st["bow_resolve_entries"], "rows_x_locations": n_frames * BOW_NQ * per_gpu,
                    "algorithmic_flops_per_unit": 2.0 * BOW_NQ * BOW_W * 128 / BOW_IMAGES,
                    # dram__bytes_read.sum + dram__bytes_write.sum of one launch at the default configuration (ncu --set full,
                    # profiles/r2_bow_tc9_final_ncu_full.md): the fp16 tiles (93 MB) once + the records written
                    "traffic": 111.4e6 if (n_frames == BOW_FRAMES and per_gpu == BOW_LOC_PER_GPU) else None,
                    "traffic_unit": "bytes per launch (ncu, profiles/r2_bow_tc9_final_ncu_full.md)"}
        cpu, checked = None, 0
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            lo = sbow.loc_bounds[0][0]
            locs = [cache[i] for i in range(lo, lo + min(4, per_gpu))]
            r2, _, _ = bow_cpu(frames, locs, 2, threads)
            n_pairs = args.cpu_sample_pairs or max(4, min(400, int(r2 * 12.0)))
            rate, dt, got = bow_cpu(frames, locs, n_pairs, threads)
            sys.path.insert(0, os.path.join(REPO, "oracle"))
            import oracle
            for f, l, words, score in got:      # scipy's own words and scores against the GPU's, near-ties excepted
                gw = None
                s = host[f, (lo + l) * BOW_IMAGES:(lo + l + 1) * BOW_IMAGES]
                want = oracle.bow_score_from_words(words, locs[l][1], locs[l][0], BOW_W)[0]
                if not np.allclose(s, want, rtol=0, atol=1e-6):
                    ow, gap = oracle.bow_words(frames[f], locs[l][2])
                    assert (gap < 1e-5).any(), "parity: frame %d location %d scores differ without a float32 near-tie" % (f, lo + l)
                checked += 1
            cpu = {"value": rate * BOW_IMAGES, "unit": "pairs/s", "cores": threads, "kind": "reference",
                   "sample": "%d (frame, location) calls = %d (frame, image) scores in %.1f s: scipy.cluster.vq.vq + sklearn normalize + "
                             "numpy dot exactly as Matcher.BOWMatch (Matcher.py:249-258); every score compared with the GPU's"
                             % (n_pairs, n_pairs * BOW_IMAGES, dt)}
        q_bytes = int(sum(f.nbytes for f in frames))
        line = {"metric": "map-image matches/sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f16", "data": "synthetic",
                "config": {"workload": bow_text(n_frames, per_gpu), "units_per_step": units_step,
                           "unit_is": "one (query frame, map image) tf-idf score; %d per (frame, location)" % BOW_IMAGES,
                           "sharding": "locations across ranks (engine.ShardedBow), float32 scores all-gathered on the device",
                           "l2": "code books per GPU %.0f MB (fp16 tiles) + %.0f MB (float32) > L2; no flush"
                                 % (per_gpu * 79 * 36864 / 1e6, per_gpu * BOW_W * 512 / 1e6)},
                "clocks": sampler.summary(),
                "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": q_bytes, "d2h_bytes_per_step": int(n_frames * total_loc * BOW_IMAGES * 4),
                        "ms_per_step": ms_e2e / args.steps, "ms_per_step_from_a_list_of_frame_arrays": ms_e2e_list / args.steps,
                        "path": "ShardedBow.score(host float32 descriptors, one pinned buffer) -> H2D -> fp16 rows -> tcgen05 -> float32 resolve -> tf-idf score -> all-gather -> D2H"},
                "gpu_launches": int(launches), "parity_pairs_checked": checked, "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    sbow.close()


def bow_reference(args, n_frames, per_gpu, total_loc):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    locs = [bow_location(i) for i in range(min(2, total_loc))]
    frames = bow_frames(2, total_loc)
    r2, _, _ = bow_cpu(frames, locs, 2, threads)
    per_step = max(2, min(200, int(r2 * min(120.0 / max(1, args.steps + args.warmup), 20.0))))
    for _ in range(args.warmup):
        bow_cpu(frames, locs, per_step, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        bow_cpu(frames, locs, per_step, threads)
    dt = time.perf_counter() - t0
    value = per_step * BOW_IMAGES * args.steps / dt
    print(json.dumps({"impl": "reference", "metric": "map-image matches/sec", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": bow_text(n_frames, per_gpu), "sample_calls_per_step": per_step},
                      "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "reference",
                                       "sample": "%d BOWMatch-equivalent calls per step (scipy vq + sklearn normalize + numpy dot)" % per_step},
                      "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# ORB
# ---------------------------------------------------------------------------------------------------------------------
def orb_workload(img_lo, img_hi, total, n_frames):
    import numpy as np
    from pymcl_b200 import synth

    def image(g):
        return synth.orb_like(np.random.default_rng(3000 + g), ORB_N)

    map_des = [image(g) for g in range(img_lo, img_hi)]
    frames = []
    for f in range(n_frames):
        rng = np.random.default_rng(8000000 + f)
        q = synth.orb_like(rng, ORB_N)
        g = (7 * f) % total
        src = map_des[g - img_lo] if img_lo <= g < img_hi else image(g)
        q[:200] = synth.orb_flip(rng, src[rng.choice(ORB_N, 200, replace=False)], 40)
        frames.append(q)
    return map_des, frames


def orb_text(n_frames, per_gpu):
    return ("ORB (BASELINE config 1's shape as a batch): %d frames x %d descriptors vs %d map images x %d descriptors per GPU, "
            "Hamming 2-NN + ratio 0.7 + count" % (n_frames, ORB_N, per_gpu, ORB_N))


def orb_cpu(frames, map_des, n_pairs, threads):
    import cv2
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import oracle
    cv2.setNumThreads(threads)
    out = []
    t0 = time.perf_counter()
    for k in range(n_pairs):
        f, i = (k // len(map_des)) % len(frames), k % len(map_des)
        out.append((f, i, oracle.cv2_orb_pair_count_knn(frames[f], map_des[i])))
    dt = time.perf_counter() - t0
    return n_pairs / dt, dt, out


def run_orb(args, H, ClockSampler, peaks):
    import numpy as np
    import torch
    import pymcl_b200  # noqa: F401
    from pymcl_b200 import _lib
    from pymcl_b200.engine import ShardedMap
    n_frames = args.frames or ORB_FRAMES
    per_gpu = args.images or ORB_IMG_PER_GPU
    if args.impl == "reference":
        return orb_reference(args, n_frames, per_gpu, per_gpu * max(1, args.gpus))
    world, rank = H.world, H.rank
    total = per_gpu * world
    ctx = _lib.Context(H.local_rank)
    lo, hi = rank * per_gpu, (rank + 1) * per_gpu
    map_des, frames = orb_workload(lo, hi, total, n_frames)
    smap = ShardedMap("ORB", map_des, ctx=ctx, local_only=True)
    queries = _lib.QueryBatch(ctx, "ORB", frames)
    algo = _lib.ALGO_TENSOR if args.algo == 1 else _lib.ALGO_SIMT
    result = {}

    def step_resident():
        result["dev"] = smap.match_counts_device(queries, algo=algo)

    packed = _lib.pack_descriptors(_lib.ORB, frames)
    pin = _lib.pinned_empty(packed[1].shape, packed[1].dtype)
    pin[...] = packed[1]
    packed_pinned = (packed[0], pin, packed[2])

    def step_e2e():                                                     # rows in one pinned buffer (like bench.py's SIFT e2e)
        result["host"] = smap.match_counts(None, algo=algo, packed=packed_pinned)

    def step_e2e_list():                                                # a Python list of per-frame arrays: + the packing pass
        result["host_list"] = smap.match_counts(frames, algo=algo)

    step_resident()
    torch.cuda.synchronize()
    c0 = result["dev"].cpu().numpy()
    step_e2e()
    assert np.array_equal(result["host"], c0), "host-buffer path and resident path disagree"
    for f in range(min(n_frames, 16)):
        assert c0[f, (7 * f) % total] >= 100, "planted matches not found: kernel output is wrong"
    sampler = ClockSampler(H.local_rank)
    sampler.start()
    time.sleep(0.3)
    for k in ("orb", "orb_fixup", "ingest"):
        ctx.profile_read(k)
    s0 = ctx.stats()
    ctx.profile_enable(True)
    l0 = ctx.launch_count()
    ms = H.timed(step_resident, args.steps, args.warmup, use_events=True)
    prof = {k: ctx.profile_read(k)[0] / (args.steps + args.warmup) for k in ("orb", "orb_fixup", "ingest")}
    ctx.profile_enable(False)
    s1 = ctx.stats()
    launches = (ctx.launch_count() - l0) * args.steps // max(1, args.steps + args.warmup)
    ms_e2e = H.timed(step_e2e, args.steps, max(1, args.warmup // 2), use_events=False)
    ms_e2e_list = H.timed(step_e2e_list, args.steps, max(1, args.warmup // 2), use_events=False)
    assert np.array_equal(result["host_list"], result["host"])
    time.sleep(0.2)
    sampler.stop()
    sampler.join(timeout=2)
    units_step = n_frames * total
    value = units_step / (ms / args.steps * 1e-3)
    e2e_value = units_step / (ms_e2e / args.steps * 1e-3)
    if rank == 0:
        ops = 2.0 * ORB_N * ORB_N * 256 * n_frames * per_gpu     # algorithmic: 256 +-1 MACs per descriptor pair
        achieved = ops / (prof["orb"] * 1e-3) / 1e12
        i8_rate, _ = ctx.microbench(0, 4000)
        popc_rate, _ = ctx.microbench(1, 20000)
        peak = i8_rate / 1e12
        roofline = {"bound": "tensor", "kernel": "tc9_kernel<ORB> (tcgen05.mma kind::i8 s8xs8->s32 128x128x32 on +-1 rows, queries in TMEM, exact keys, "
                                                 "top-2 group epilogue)" if args.algo == 1 else "orb_knn_kernel (xor + popc)",
                    "achieved": achieved, "peak": peak, "unit": "TOP/s", "frac": achieved / peak,
                    "peak_source": "int8 tensor pipe measured live in this run (mcl_microbench); MEASURED_PEAKS.json carries no int8 figure",
                    "frac_of_nominal": achieved / 4500.0,
                    "popc_roofline_pairs_per_s": popc_rate / (ORB_N * ORB_N * 8.0),
                    "times_the_popc_roofline": (n_frames * per_gpu / (prof["orb"] * 1e-3)) / (popc_rate / (ORB_N * ORB_N * 8.0)),
                    "kernel_ms_per_step": prof["orb"], "kernel_share_of_step": prof["orb"] / (ms / args.steps),
                    "fixup_ms_per_step": prof["orb_fixup"], "prep_ms_per_step": prof["ingest"],
                    "candidates_per_row_image": (s1["orb_candidates"] - s0["orb_candidates"]) / (args.steps + args.warmup) / (n_frames * ORB_N * per_gpu),
                    "repaired_pairs_per_step": (s1["repaired_pairs"] - s0["repaired_pairs"]) / (args.steps + args.warmup),
                    "algorithmic_ops_per_unit": 2.0 * ORB_N * ORB_N * 256,
                    # dram__bytes_read.sum + dram__bytes_write.sum of one launch at the default configuration (ncu --set full,
                    # profiles/r2_orb_tc9_final_ncu_full.md): map tiles 52 MB + query rows 26 MB once, the rest from L2
                    "traffic": 83.2e6 if (n_frames == ORB_FRAMES and per_gpu == ORB_IMG_PER_GPU and args.algo == 1 and ORB_N == 500) else None,
                    "traffic_unit": "bytes per launch (ncu, profiles/r2_orb_tc9_final_ncu_full.md)"}
        cpu, checked = None, 0
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            r4, _, _ = orb_cpu(frames, map_des, 8, threads)
            n_pairs = args.cpu_sample_pairs or max(16, min(8000, int(r4 * 10.0)))
            rate, dt, got = orb_cpu(frames, map_des, n_pairs, threads)
            for f, i, want in got:
                assert int(c0[f, lo + i]) == want, "parity: frame %d image %d: GPU %d, cv2 %d" % (f, lo + i, int(c0[f, lo + i]), want)
            checked = len(got)
            cpu = {"value": rate, "unit": "pairs/s", "cores": threads, "kind": "port",
                   "sample": "%d (frame, image) pairs of %dx%d x 256 bit in %.1f s: cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2) + ratio lambda "
                             "(oracle.cv2_orb_pair_count_knn); every count compared with the GPU's" % (n_pairs, ORB_N, ORB_N, dt)}
        line = {"metric": "map-image matches/sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "s8", "data": "synthetic",
                "config": {"workload": orb_text(n_frames, per_gpu), "units_per_step": units_step,
                           "sharding": "map images across ranks (engine.ShardedMap), int32 counts all-gathered on the device",
                           "l2": "inputs per step %.0f MB (+-1 byte tiles %.0f MB + queries %.0f MB); fits L2: the kernel is tensor bound, not memory bound"
                                 % ((per_gpu * 512 + n_frames * 512) * 256 / 1e6, per_gpu * 512 * 256 / 1e6, n_frames * 512 * 256 / 1e6)},
                "clocks": sampler.summary(),
                "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": int(n_frames * ORB_N * 32), "d2h_bytes_per_step": int(n_frames * total * 4),
                        "ms_per_step": ms_e2e / args.steps, "ms_per_step_from_a_list_of_frame_arrays": ms_e2e_list / args.steps,
                        "path": "ShardedMap.match_counts(host uint8 descriptors, one pinned buffer) -> H2D -> +-1 rows -> tcgen05 -> popc fix-up -> all-gather -> D2H"},
                "gpu_launches": int(launches), "parity_pairs_checked": checked, "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    smap.close()


def orb_reference(args, n_frames, per_gpu, total):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    map_des, frames = orb_workload(0, 8, max(8, total), 2)
    r, _, _ = orb_cpu(frames, map_des, 8, threads)
    per_step = max(8, min(4000, int(r * min(120.0 / max(1, args.steps + args.warmup), 20.0))))
    for _ in range(args.warmup):
        orb_cpu(frames, map_des, per_step, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orb_cpu(frames, map_des, per_step, threads)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    print(json.dumps({"impl": "reference", "metric": "map-image matches/sec", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                      "config": {"workload": orb_text(n_frames, per_gpu), "sample_pairs_per_step": per_step},
                      "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port",
                                       "sample": "%d pairs per step, cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2) + ratio lambda" % per_step},
                      "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
