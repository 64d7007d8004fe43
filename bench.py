#!/usr/bin/env python
"""bench.py — map-image matches/sec of the descriptor-matching hot path (BASELINE.json), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload sift|bow|orb]
                    [--scaling weak|strong] [--frames F] [--images I]

--workload sift (default, BASELINE config 3, the configuration the metric is quoted on): 64 locations x 50 images x 2048 SIFT
  descriptors sharded over 8 GPUs = 8 locations (400 images, 105 MB resident) per GPU; one step = one batch of 200 query frames
  x 2048 descriptors matched against every image (exact 2-NN + Lowe ratio test + good-match count), the int32 counts
  all-gathered.  A unit is one (query frame, map image) pair.  --scaling weak keeps 400 images per GPU (N = 8 is config 3
  itself); --scaling strong keeps config 3's 3200 images in total and splits them over the N ranks.
--workload bow (BASELINE config 5): 256 locations x 100 images, 10 000 visual words per location, sharded by location
  (32 per GPU, weak); one step = 50 frames x 300 descriptors assigned to words and scored against every location;
  a unit is one (query frame, map image) tf-idf score (bench_other.py).
--workload orb (BASELINE config 1's shape as a batch): 200 frames x 400 images x 500 ORB descriptors.

  value         inputs resident in HBM: the product's sharded path (engine.ShardedMap.match_counts_device: this rank's kernels +
                the device all-gather of the counts), CUDA events on the launching stream, max over ranks
  e2e           the call a user makes with HOST descriptors (ShardedMap.match_counts on pinned float32 rows as cv2 hands them
                over): H2D -> ingest -> kernels -> all-gather -> D2H of the counts, wall clock, max over ranks
  roofline      the dominant kernel's algorithmic ops / its own CUDA-event time, against the tensor-pipe peak measured live
                by mcl_microbench (MEASURED_PEAKS.json carries no int8 figure; 2 x its bf16 numbers are printed beside it)
  cpu_baseline  cv2.BFMatcher.knnMatch(k=2) + the reference's ratio lambda (oracle.cv2_l2_pair_count) on the box's host
                cores, on a bounded sample of the SAME (frame, image) pairs -- whose counts are then compared with the GPU's
                (parity_pairs_checked)
  distributions the same kernels on other descriptor statistics (cv2-SIFT rows of textured 800x600 images, uniform random
                bytes): pairs/s, candidates per (row, image), fix-up share, exact rescans, repaired pairs
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

NQ = 2048          # descriptors per query frame (SIFT 800x600)
NM = 2048          # descriptors per map image
LOC_PER_GPU = 8
IMG_PER_LOC = 50
CFG3_IMAGES = 64 * IMG_PER_LOC
FRAMES = 200
OPS_PER_UNIT = 2.0 * NQ * NM * 128   # algorithmic int-ops of one (frame, image) pair: Nq*Nm*128 MACs


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sift", choices=["sift", "bow", "orb"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--images", type=int, default=0, help="map images per GPU (weak) / in total (strong); BOW: locations")
    ap.add_argument("--cpu-sample-pairs", type=int, default=0, help="0 = auto (about 10-20 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-distributions", action="store_true")
    ap.add_argument("--algo", type=int, default=1, help="1 = tensor cores (product), 2 = SIMT")
    ap.add_argument("--opt1", type=int, default=0, help="libmcl diagnostic option 1")
    return ap.parse_args()


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return {"bf16_burst": d.get("bf16_tflops"), "bf16_sustained": d.get("bf16_tflops_sustained"),
                "hbm": d.get("hbm_gbs"), "src": "measured"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "src": "fallback"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# SIFT workload (BASELINE config 3)
# ---------------------------------------------------------------------------------------------------------------------
def sift_workload(img_lo, img_hi, total_images, n_frames):
    """Seeded synthetic descriptors of BASELINE cfg3's shape.  Global map image g is generated from its own seed, so a rank
    holds the same image whatever the sharding; the query batch is the same on every rank.  Frame f re-observes 25 % of
    the rows of global map image (7 f mod total) -- regenerated from that image's seed when another rank holds it."""
    import numpy as np
    from pymcl_b200 import synth

    def image(g):
        return synth.sift_like(np.random.default_rng(1000 + g), NM)

    map_des = [image(g) for g in range(img_lo, img_hi)]
    frames = []
    for f in range(n_frames):
        rng = np.random.default_rng(7000000 + f)
        q = synth.sift_like(rng, NQ)
        g = (7 * f) % total_images
        src = map_des[g - img_lo] if img_lo <= g < img_hi else image(g)
        rows = rng.choice(NM, size=NQ // 4, replace=False)
        dst = rng.choice(NQ, size=NQ // 4, replace=False)
        q[dst] = synth.sift_noisy_copy(rng, src[rows], 4.0)
        frames.append(q)
    return map_des, frames


def cpu_pairs(frames, map_des, n_pairs, threads):
    """cv2.BFMatcher(NORM_L2).knnMatch(k=2) + the reference's ratio lambda (Matcher.py:279-280) + len(), over the pairs
    (frame 0, image 0), (frame 0, image 1), ...  Returns (pairs per second, seconds, [(frame, image, count)])."""
    import cv2
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import oracle
    cv2.setNumThreads(threads)
    oracle.cv2_l2_pair_count(frames[0], map_des[0])  # warm
    out = []
    t0 = time.perf_counter()
    for k in range(n_pairs):
        f, i = (k // len(map_des)) % len(frames), k % len(map_des)
        out.append((f, i, oracle.cv2_l2_pair_count(frames[f], map_des[i])))
    dt = time.perf_counter() - t0
    return n_pairs / dt, dt, out


def workload_text(args, n_frames, per_gpu, total):
    return ("SIFT 800x600 whole-map matching (BASELINE config 3): %d frames x %d descriptors vs %d map images x %d descriptors "
            "per GPU (%d locations x %d images), 2-NN + ratio 0.7 + count" % (n_frames, NQ, per_gpu, NM, max(1, per_gpu // IMG_PER_LOC), IMG_PER_LOC))


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path (cv2 brute-force knnMatch + ratio lambda; the oracle's
    cv2-backed port, the reference itself being Python that cannot travel to the GPU box), on a bounded sample per step,
    all host threads.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import cv2
    threads = os.cpu_count() or 1
    n_frames = args.frames or FRAMES
    per_gpu = args.images or LOC_PER_GPU * IMG_PER_LOC
    if args.scaling == "strong":
        per_gpu = (args.images or CFG3_IMAGES) // max(1, args.gpus)
    map_des, frames = sift_workload(0, 8, max(8, per_gpu * max(1, args.gpus)), 2)
    rate, _, _ = cpu_pairs(frames, map_des, 4, threads)   # calibrate the per-step sample: steps+warmup within a few minutes
    budget_s = 120.0 / max(1, args.steps + args.warmup)
    per_step = max(2, min(400, int(rate * min(budget_s, 20.0))))
    for _ in range(args.warmup):
        cpu_pairs(frames, map_des, per_step, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_pairs(frames, map_des, per_step, threads)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": "map-image matches/sec", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(args, n_frames, per_gpu, per_gpu * max(1, args.gpus)),
                   "sample_pairs_per_step": per_step, "cv2": cv2.__version__},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port",
                         "sample": "%d (frame,image) pairs of %dx%dx128 per step, cv2.BFMatcher(NORM_L2).knnMatch(k=2)"
                                   " + ratio lambda, cv2.setNumThreads(%d)" % (per_step, NQ, NM, threads)},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class Harness:
    """process-group set-up, barriers, device / wall-clock timing with max over ranks."""

    def __init__(self, args):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, self.world))
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        # a real (non-default) stream: the library launches on it, torch's events and NCCL see the same stream
        self.stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.stream)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, warmup, use_events):
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        if use_events:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            for _ in range(steps):
                fn()
            e1.record(self.stream)
            self.barrier()
            ms = e0.elapsed_time(e1)
        else:
            t0 = time.perf_counter()
            for _ in range(steps):
                fn()
            self.barrier()
            ms = 1e3 * (time.perf_counter() - t0)
        if self.world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def finish(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def sift_distributions(ctx, _lib, np):
    """The SIFT kernels on other descriptor statistics than the synthetic generator's: how often the bracketed pre-filter
    has to hand a (row, image) pair to the exact fix-up, what that costs, and whether anything falls off the fast path."""
    import cv2
    from pymcl_b200 import synth
    out = []

    def measure(tag, map_des, frames, note):
        m = _lib.MapIndex(ctx, "SIFT", map_des)
        q = _lib.QueryBatch(ctx, "SIFT", frames)
        import torch
        d = torch.zeros((len(frames), len(map_des)), dtype=torch.int32, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        run = lambda: m.match_counts_device(q, d.data_ptr(), algo=_lib.ALGO_TENSOR, stream=st)
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        s0 = ctx.stats()
        for k in _lib.PROFILE_SLOTS:
            ctx.profile_read(k)
        ctx.profile_enable(True)
        reps = 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        tc, _ = ctx.profile_read("sift_tc")
        fix, _ = ctx.profile_read("sift_fixup")
        simt, _ = ctx.profile_read("sift_simt")
        ctx.profile_enable(False)
        s1 = ctx.stats()
        rows_q = sum(len(f) for f in frames)
        rows_m = sum(len(x) for x in map_des)
        pairs = len(frames) * len(map_des)
        ops = 2.0 * rows_q * rows_m * 128
        # spot check against the exact SIMT kernel (itself pinned to the oracle in tests/)
        want = m.match_counts(frames[:2], algo=_lib.ALGO_SIMT)
        got = d.cpu().numpy()[:2]
        out.append({"input": tag, "note": note, "frames": len(frames), "images": len(map_des),
                    "rows_per_frame": rows_q / len(frames), "rows_per_image": rows_m / len(map_des),
                    "pairs_per_s": pairs / (ms * 1e-3), "ms": ms, "tensor_kernel_TOPs": ops / (tc / reps * 1e-3) / 1e12,
                    "candidates_per_row_image": (s1["sift_candidates"] - s0["sift_candidates"]) / reps / max(1, rows_q * len(map_des)),
                    "good_matches_per_row_image": float(d.sum().item()) / max(1, rows_q * len(map_des)),
                    "fixup_share": fix / max(1e-9, tc + fix + simt), "repair_share": simt / max(1e-9, tc + fix + simt),
                    "fixup_ms": fix / reps,
                    "exact_rescans": (s1["sift_rescans"] - s0["sift_rescans"]) // reps,
                    "repaired_pairs": (s1["repaired_pairs"] - s0["repaired_pairs"]) // reps,
                    "equals_exact_kernel": bool(np.array_equal(want, got))})
        q.close()
        m.close()

    # (i) real cv2-SIFT descriptors: textured 800x600 views of a synthetic panorama, neighbouring views overlap
    views, queries = synth.textured_views(5, n_views=48, n_queries=12, view=(800, 600))
    sift = cv2.SIFT_create()
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as pool:
        des = list(pool.map(lambda im: cv2.SIFT_create().detectAndCompute(im, None)[1], views + queries))
    measure("cv2_sift_800x600", des[:len(views)], des[len(views):], "cv2.SIFT_create() on textured 800x600 synthetic views")
    # (ii) uniform random bytes: no structure at all, norms spread over +-8 %, every distance nearly the same
    rng = np.random.default_rng(17)
    measure("uniform_u8", [rng.integers(0, 256, size=(NM, 128)).astype(np.uint8) for _ in range(100)],
            [rng.integers(0, 256, size=(NQ, 128)).astype(np.uint8) for _ in range(24)], "uniform random bytes, 2048 rows")
    # (iii) the bench generator itself at the same reduced size, for reference
    md, fr = sift_workload(0, 100, 100, 24)
    measure("synthetic_sift_like", md, fr, "round(512 g/|g|), g ~ Gamma(0.6): the timed workload's generator")
    return out


def run_sift(args, H):
    import numpy as np
    import torch
    import pymcl_b200  # noqa: F401
    from pymcl_b200 import _lib
    from pymcl_b200.engine import ShardedMap

    world, rank = H.world, H.rank
    ctx = _lib.Context(H.local_rank)  # fails loudly without libmcl.so / an sm_100 GPU
    info = ctx.info()
    if args.opt1:
        ctx.set_option(1, args.opt1)
    n_frames = args.frames or FRAMES
    if args.scaling == "weak":
        per_gpu = args.images or LOC_PER_GPU * IMG_PER_LOC
        total = per_gpu * world
        lo, hi = rank * per_gpu, (rank + 1) * per_gpu
    else:
        total = args.images or CFG3_IMAGES
        lo, hi = total * rank // world, total * (rank + 1) // world
        per_gpu = hi - lo
    map_des, frames = sift_workload(lo, hi, total, n_frames)
    smap = ShardedMap("SIFT", map_des, ctx=ctx, local_only=True)        # the product's sharded index
    assert smap.n_images == total and (smap.lo, smap.hi) == (lo, hi)
    packed = _lib.pack_descriptors(_lib.SIFT, frames)                   # float32, as cv2.SIFT hands them over
    pin = _lib.pinned_empty(packed[1].shape, np.float32)
    pin[...] = packed[1]
    packed_pinned = (packed[0], pin, packed[2])
    queries = _lib.QueryBatch(ctx, "SIFT", packed=packed_pinned)        # resident copy for `value`
    algo = args.algo
    result = {}

    def step_resident():
        result["dev"] = smap.match_counts_device(queries, algo=algo)    # kernels + device all-gather of the counts

    def step_e2e():
        result["host"] = smap.match_counts(None, algo=algo, packed=packed_pinned)

    # ---- correctness guards on the timed configuration: the planted image must win, and both paths must agree
    step_resident()
    torch.cuda.synchronize()
    c0 = result["dev"].cpu().numpy()
    step_e2e()
    assert np.array_equal(result["host"], c0), "host-buffer path and resident path disagree"
    for f in range(min(n_frames, 16)):
        g = (7 * f) % total
        assert c0[f, g] >= 0.9 * (NQ // 4), "planted matches not found: kernel output is wrong"

    # ---- value: resident inputs, device timed
    sampler = ClockSampler(H.local_rank)
    sampler.start()
    time.sleep(0.3)
    s0 = ctx.stats()
    ctx.profile_enable(True)
    l0 = ctx.launch_count()
    ms = H.timed(step_resident, args.steps, args.warmup, use_events=True)
    tc_ms, tc_n = ctx.profile_read("sift_tc")     # profile events cover warm-up + timed steps of this rank's kernels
    fix_ms, fix_n = ctx.profile_read("sift_fixup")
    simt_ms, simt_n = ctx.profile_read("sift_simt")
    ctx.profile_enable(False)
    s1 = ctx.stats()
    launches = (ctx.launch_count() - l0) * args.steps // max(1, args.steps + args.warmup)
    # ---- e2e: host buffers through the public sharded call, wall clock around H2D + kernels + all-gather + D2H
    ms_e2e = H.timed(step_e2e, args.steps, max(1, args.warmup // 2), use_events=False)
    time.sleep(0.2)
    sampler.stop()
    sampler.join(timeout=2)
    clocks = sampler.summary()

    units_step = n_frames * total
    value = units_step / (ms / args.steps * 1e-3)
    e2e_value = units_step / (ms_e2e / args.steps * 1e-3)
    if rank == 0:
        pk = peaks()
        steps_prof = args.steps + args.warmup
        tc_ms_step = tc_ms / steps_prof
        ops_step = OPS_PER_UNIT * n_frames * per_gpu      # this rank's tcgen05 launches of one step
        achieved = ops_step / (tc_ms_step * 1e-3) / 1e12
        i8_rate, _ = ctx.microbench(0, 4000)
        peak = i8_rate / 1e12
        default_cfg = n_frames == FRAMES and per_gpu == LOC_PER_GPU * IMG_PER_LOC and algo == 1
        roofline = {
            "bound": "tensor", "kernel": "sift_tc4_kernel (tcgen05.mma kind::i8 u8xu8->s32 128x128x32, queries in TMEM, norm-sorted map tiles, fused bracketed top-2 epilogue)",
            "achieved": achieved, "peak": peak, "unit": "TOP/s", "frac": achieved / peak,
            "peak_source": "int8 tensor pipe measured live in this run (mcl_microbench: back-to-back tcgen05.mma kind::i8 "
                           "128x128x32 on all SMs); MEASURED_PEAKS.json (%s) carries no int8 figure" % pk["src"],
            "peak_2x_bf16_sustained": 2.0 * pk["bf16_sustained"], "frac_of_2x_bf16_sustained": achieved / (2.0 * pk["bf16_sustained"]),
            "peak_2x_bf16_burst": 2.0 * pk["bf16_burst"], "frac_of_2x_bf16_burst": achieved / (2.0 * pk["bf16_burst"]),
            "peak_nominal": 4500.0, "frac_of_nominal": achieved / 4500.0,
            "kernel_ms_per_step": tc_ms_step, "kernel_launches_per_step": tc_n / steps_prof,
            "kernel_share_of_step": tc_ms_step / (ms / args.steps),
            "fixup_ms_per_step": fix_ms / steps_prof, "repair_ms_per_step": simt_ms / steps_prof,
            "candidates_per_row_image": (s1["sift_candidates"] - s0["sift_candidates"]) / steps_prof / (n_frames * NQ * per_gpu),
            "exact_rescans_per_step": (s1["sift_rescans"] - s0["sift_rescans"]) / steps_prof,
            "repaired_pairs_per_step": (s1["repaired_pairs"] - s0["repaired_pairs"]) / steps_prof,
            "algorithmic_ops_per_unit": OPS_PER_UNIT, "units_per_step_per_gpu": n_frames * per_gpu,
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch at the default configuration, from the ncu --set full
            # capture summarised in profiles/r1_sift_tc4_ncu_full.md (tensor-bound kernel: 0.2 % of the HBM peak)
            "traffic": 290.0e6 if default_cfg else None,
            "traffic_unit": "bytes per launch (ncu, profiles/r1_sift_tc4_ncu_full.md)",
        }
        cpu = None
        parity_checked = 0
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            r4, _, _ = cpu_pairs(frames, map_des, 4, threads)
            n_pairs = args.cpu_sample_pairs or max(8, min(2000, int(r4 * 15.0)))
            rate, dt, counts = cpu_pairs(frames, map_des, n_pairs, threads)
            # the CPU sample is the reference's arithmetic on pairs of the timed workload: their counts must equal the GPU's
            for f, i, want in counts:
                got = int(c0[f, lo + i])
                assert got == want, "parity: frame %d image %d: GPU %d, cv2 %d" % (f, lo + i, got, want)
            parity_checked = len(counts)
            cpu = {"value": rate, "unit": "pairs/s", "cores": threads, "kind": "port",
                   "sample": "%d (frame,image) pairs of %dx%dx128 of this workload in %.1f s: cv2.BFMatcher(NORM_L2)"
                             ".knnMatch(k=2) + ratio lambda (oracle.cv2_l2_pair_count), cv2.setNumThreads(%d); every "
                             "count compared with the GPU's" % (n_pairs, NQ, NM, dt, threads)}
        dists = None
        if not args.no_distributions and world == 1:
            dists = sift_distributions(ctx, _lib, np)
        h2d = int(packed_pinned[1].nbytes // world + packed_pinned[0].nbytes) if world > 1 else int(packed_pinned[1].nbytes + packed_pinned[0].nbytes)
        d2h = int(n_frames * total * 4)
        line = {
            "metric": "map-image matches/sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_text(args, n_frames, per_gpu, total),
                       "units_per_step": units_step, "total_images": total,
                       "sharding": "map images across ranks (engine.ShardedMap); query rows uploaded once per box and all-gathered "
                                   "over NVLink; int32 counts all-gathered on the device",
                       "l2": "inputs per step %.0f MB (map %.0f MB + queries %.0f MB) > L2 %.0f MB; no flush"
                             % ((per_gpu * NM + n_frames * NQ) * 128 / 1e6, per_gpu * NM * 128 / 1e6,
                                n_frames * NQ * 128 / 1e6, info["l2_bytes"] / 1e6),
                       "sm_count": info["sm_count"]},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps,
                    "path": "ShardedMap.match_counts(host float32 descriptors, pinned) -> H2D (each frame on one rank) -> NVLink "
                            "all-gather of the rows -> ingest -> tcgen05 -> all-gather of counts -> D2H" if world > 1 else
                            "ShardedMap.match_counts = mcl_match_counts(host float32 descriptors, pinned) -> H2D -> ingest -> tcgen05 -> D2H counts"},
            "gpu_launches": int(launches),
            "parity_pairs_checked": parity_checked,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "distributions": dists,
        }
        print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        if args.workload == "sift":
            run_reference(args)
        else:
            import bench_other
            getattr(bench_other, "run_" + args.workload)(args, None, ClockSampler, peaks)
        return
    H = Harness(args)
    if args.workload == "sift":
        run_sift(args, H)
    else:
        import bench_other
        getattr(bench_other, "run_" + args.workload)(args, H, ClockSampler, peaks)
    H.finish()


if __name__ == "__main__":
    main()
