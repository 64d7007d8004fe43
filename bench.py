#!/usr/bin/env python
"""bench.py — map-image matches/sec of the SIFT whole-map matching hot path (BASELINE.json config 3).

Workload (per GPU, weak scaling): 8 locations x 50 images x 2048 SIFT descriptors resident in HBM (105 MB),
one step = one batch of 200 query frames x 2048 descriptors (BASELINE config 3 = 64 locations over 8 GPUs)
matched against this rank's slice, then ONE all-gather of the int32 per-image counts.  A unit is one
(query frame, map image) pair taken through exact 2-NN + Lowe ratio test + good-match count.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--frames F] [--images I]

  value      resident inputs: mcl_match_counts_device on HBM-resident queries, CUDA events, max over ranks
  e2e        the host-buffer C-ABI call mcl_match_counts (pinned float32 descriptors as cv2 hands them over ->
             H2D -> ingest -> kernels -> D2H of counts) + the same all-gather
  roofline   sift_tc_kernel (tcgen05 u8xu8->s32): algorithmic int-ops / CUDA-event time of that kernel alone,
             against 2 x the measured bf16 tensor peak (MEASURED_PEAKS.json; int8 runs at twice the bf16 rate),
             with the live tcgen05 kind::i8 issue-rate microbenchmark and the 4.5 POP/s nominal beside it
  cpu_baseline  oracle (cv2.BFMatcher.knnMatch(k=2) + the reference's ratio lambda) on the box's host cores
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
FRAMES = 200
OPS_PER_UNIT = 2.0 * NQ * NM * 128   # algorithmic int-ops of one (frame, image) pair: Nq*Nm*128 MACs


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES)
    ap.add_argument("--images", type=int, default=LOC_PER_GPU * IMG_PER_LOC, help="map images per GPU")
    ap.add_argument("--cpu-sample-pairs", type=int, default=0, help="0 = auto (about 10-20 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--algo", type=int, default=1, help="1 = sift_tc4 (product), 5 = sift_tc3, 4 = sift_tc2, 3 = sift_tc, 2 = SIMT")
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


def make_workload(rank, n_images, n_frames):
    """Seeded synthetic descriptors of BASELINE cfg3's shape.  The map slice depends on the rank, the query batch
    is the same on every rank.  Frame f re-observes 25% of the rows of global map image (7f mod total)."""
    import numpy as np
    from pymcl_b200 import synth
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rng_q = np.random.default_rng(3)
    frames = [synth.sift_like(rng_q, NQ) for _ in range(n_frames)]
    rng_m = np.random.default_rng(1000 + rank)
    map_des = [synth.sift_like(rng_m, NM) for _ in range(n_images)]
    total = n_images * world
    for f in range(n_frames):
        g = (7 * f) % total
        if g // n_images == rank:
            src = map_des[g % n_images]
            rows = rng_q.choice(NM, size=NQ // 4, replace=False)
            dst = rng_q.choice(NQ, size=NQ // 4, replace=False)
            frames[f][dst] = synth.sift_noisy_copy(np.random.default_rng(f), src[rows], 4.0)
        else:  # keep the query rng stream identical on every rank
            rng_q.choice(NM, size=NQ // 4, replace=False)
            rng_q.choice(NQ, size=NQ // 4, replace=False)
    return map_des, frames


def cpu_pairs_per_s(frames, map_des, n_pairs, threads):
    """cv2.BFMatcher(NORM_L2).knnMatch(k=2) + the reference's ratio lambda (Matcher.py:279-280) + len()."""
    import cv2
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import oracle
    cv2.setNumThreads(threads)
    oracle.cv2_l2_pair_count(frames[0], map_des[0])  # warm
    t0 = time.perf_counter()
    done = 0
    f = i = 0
    while done < n_pairs:
        oracle.cv2_l2_pair_count(frames[f % len(frames)], map_des[i % len(map_des)])
        done += 1
        i += 1
        if i % len(map_des) == 0:
            f += 1
    dt = time.perf_counter() - t0
    return done / dt, dt


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path (cv2 brute-force knnMatch + ratio lambda;
    the oracle's cv2-backed port, the reference itself being Python that cannot travel to the GPU box), on a bounded
    sample per step, all host threads.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import cv2
    threads = os.cpu_count() or 1
    n_img = min(args.images, 8)
    map_des, frames = make_workload(0, n_img, 2)
    # size the per-step sample so that steps+warmup finish within a few minutes: calibrate on 4 pairs
    rate, _ = cpu_pairs_per_s(frames, map_des, 4, threads)
    budget_s = 120.0 / max(1, args.steps + args.warmup)
    per_step = max(2, min(400, int(rate * min(budget_s, 20.0))))
    for _ in range(args.warmup):
        cpu_pairs_per_s(frames, map_des, per_step, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_pairs_per_s(frames, map_des, per_step, threads)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": "map-image matches/sec", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "SIFT 800x600 whole-map matching (BASELINE config 3): %d frames x %d descriptors vs "
                               "%d map images x %d descriptors per GPU (%d locations x %d images), 2-NN + ratio 0.7 + count"
                               % (args.frames, NQ, args.images, NM, max(1, args.images // IMG_PER_LOC), IMG_PER_LOC),
                   "sample_pairs_per_step": per_step, "cv2": cv2.__version__},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port",
                         "sample": "%d (frame,image) pairs of %dx%dx128 per step, cv2.BFMatcher(NORM_L2).knnMatch(k=2)"
                                   " + ratio lambda, cv2.setNumThreads(%d)" % (per_step, NQ, NM, threads)},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import numpy as np
    import torch
    import pymcl_b200  # noqa: F401
    from pymcl_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = _lib.Context(local_rank)  # fails loudly without libmcl.so / an sm_100 GPU
    info = ctx.info()
    if args.opt1:
        ctx.set_option(1, args.opt1)

    n_images, n_frames = args.images, args.frames
    map_des, frames = make_workload(rank, n_images, n_frames)
    resident = _lib.MapIndex(ctx, "SIFT", map_des)
    packed = _lib.pack_descriptors(_lib.SIFT, frames)              # float32, as cv2.SIFT hands them over
    pin = _lib.pinned_empty(packed[1].shape, np.float32)
    pin[...] = packed[1]
    packed_pinned = (packed[0], pin, packed[2])
    queries = _lib.QueryBatch(ctx, "SIFT", packed=packed_pinned)   # resident copy for `value`

    # a real (non-default) stream: the library launches on it, torch's events and NCCL see the same stream
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    d_counts = torch.zeros((n_frames, n_images), dtype=torch.int32, device="cuda")
    d_all = torch.zeros((world, n_frames, n_images), dtype=torch.int32, device="cuda") if world > 1 else None

    def step_resident():
        resident.match_counts_device(queries, d_counts.data_ptr(), algo=args.algo, stream=stream.cuda_stream)
        if world > 1:
            dist.all_gather_into_tensor(d_all, d_counts)

    def step_e2e():
        c = resident.match_counts(None, algo=args.algo, packed=packed_pinned)
        if world > 1:
            t = torch.from_numpy(c).cuda(non_blocking=True)
            dist.all_gather_into_tensor(d_all, t)
            return d_all.cpu().numpy()
        return c

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, use_events):
        for _ in range(warmup):
            fn()
        barrier()
        if use_events:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
            barrier()
            ms = e0.elapsed_time(e1)
        else:
            t0 = time.perf_counter()
            for _ in range(steps):
                fn()
            barrier()
            ms = 1e3 * (time.perf_counter() - t0)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- correctness guard on the timed configuration (planted image must win; nothing is skipped)
    step_resident()
    torch.cuda.synchronize()
    c0 = d_counts.cpu().numpy()
    total_images = n_images * world
    for f in range(min(n_frames, 8)):
        g = (7 * f) % total_images
        if g // n_images == rank:
            assert c0[f, g % n_images] >= 0.9 * (NQ // 4), "planted matches not found: kernel output is wrong"

    # ---- value: resident inputs, device timed
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    ctx.profile_enable(True)
    l0 = ctx.launch_count()
    ms = timed(step_resident, args.steps, args.warmup, use_events=True)
    # profile events cover warm-up + timed steps of this rank's kernels
    tc_ms, tc_n = ctx.profile_read("sift_tc")
    fix_ms, fix_n = ctx.profile_read("sift_fixup")
    simt_ms, simt_n = ctx.profile_read("sift_simt")
    ctx.profile_enable(False)
    launches = (ctx.launch_count() - l0) * args.steps // max(1, args.steps + args.warmup)
    # ---- e2e: host buffers through the C-ABI, wall clock around H2D + kernels + D2H (+ all-gather)
    ms_e2e = timed(step_e2e, args.steps, max(1, args.warmup // 2), use_events=False)
    time.sleep(0.2)
    sampler.stop()
    sampler.join(timeout=2)
    clocks = sampler.summary()

    units_step = n_frames * n_images * world
    value = units_step / (ms / args.steps * 1e-3)
    e2e_value = units_step / (ms_e2e / args.steps * 1e-3)

    if rank == 0:
        pk = peaks()
        # dominant kernel: all tcgen05 launches of one step process n_frames*n_images units on this rank
        steps_prof = args.steps + args.warmup
        tc_ms_step = tc_ms / steps_prof
        ops_step = OPS_PER_UNIT * n_frames * n_images
        achieved = ops_step / (tc_ms_step * 1e-3) / 1e12
        i8_rate, _ = ctx.microbench(0, 4000)
        peak = i8_rate / 1e12
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
            "algorithmic_ops_per_unit": OPS_PER_UNIT, "units_per_step_per_gpu": n_frames * n_images,
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch at the default configuration, from the ncu --set full
            # capture summarised in profiles/r1_sift_tc4_ncu_full.md (tensor-bound kernel: 0.2 % of the HBM peak)
            "traffic": 290.0e6 if (n_frames == FRAMES and n_images == LOC_PER_GPU * IMG_PER_LOC and args.algo == 1) else None,
            "traffic_unit": "bytes per launch (ncu, profiles/r1_sift_tc4_ncu_full.md)",
        }
        cpu = None
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            r4, _ = cpu_pairs_per_s(frames, map_des, 4, threads)
            n_pairs = args.cpu_sample_pairs or max(8, min(2000, int(r4 * 15.0)))
            rate, dt = cpu_pairs_per_s(frames, map_des, n_pairs, threads)
            cpu = {"value": rate, "unit": "pairs/s", "cores": threads, "kind": "port",
                   "sample": "%d (frame,image) pairs of %dx%dx128 of this workload in %.1f s: cv2.BFMatcher(NORM_L2)"
                             ".knnMatch(k=2) + ratio lambda (oracle.cv2_l2_pair_count), cv2.setNumThreads(%d)"
                             % (n_pairs, NQ, NM, dt, threads)}
        h2d = int(packed_pinned[1].nbytes + packed_pinned[0].nbytes)
        d2h = int(n_frames * n_images * 4)
        line = {
            "metric": "map-image matches/sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "SIFT 800x600 whole-map matching (BASELINE config 3): %d frames x %d descriptors vs "
                                   "%d map images x %d descriptors per GPU (%d locations x %d images), 2-NN + ratio 0.7 + count"
                                   % (n_frames, NQ, n_images, NM, max(1, n_images // IMG_PER_LOC), IMG_PER_LOC),
                       "units_per_step": units_step, "sharding": "map images across ranks, int32 counts all-gathered",
                       "l2": "inputs per step %.0f MB (map %.0f MB + queries %.0f MB) > L2 %.0f MB; no flush"
                             % ((n_images * NM + n_frames * NQ) * 128 / 1e6, n_images * NM * 128 / 1e6,
                                n_frames * NQ * 128 / 1e6, info["l2_bytes"] / 1e6),
                       "sm_count": info["sm_count"]},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps,
                    "path": "mcl_match_counts(host float32 descriptors, pinned) -> H2D -> ingest -> tcgen05 -> D2H counts"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
