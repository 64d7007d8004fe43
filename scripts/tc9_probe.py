"""Development probe of the tc9 kernels (ORB on the tensor cores, BOW fp16 assignment): kernel times through the library's
own CUDA-event profile slots, fraction of the tensor peak, candidates / rescans.  python scripts/tc9_probe.py [orb] [bow]"""
import json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import pymcl_b200
from pymcl_b200 import _lib, synth

ctx = _lib.Context(0)
names = sys.argv[1:] or ["orb", "bow"]
i8_peak, _ = ctx.microbench(0, 4000)


def prof(fn, slots, reps=5, warm=2):
    for _ in range(warm):
        fn()
    for k in _lib.PROFILE_SLOTS:
        ctx.profile_read(k)
    ctx.profile_enable(True)
    for _ in range(reps):
        fn()
    r = {k: ctx.profile_read(k)[0] / reps for k in slots}
    ctx.profile_enable(False)
    return r


if "orb" in names:
    import torch
    rng = np.random.default_rng(1)
    for F, I in ((200, 400), (1, 50)):
        map_des = [synth.orb_like(rng, 500) for _ in range(I)]
        frames = [synth.orb_like(rng, 500) for _ in range(F)]
        for f in range(F):
            frames[f][:150] = synth.orb_flip(rng, map_des[(7 * f) % I][:150], 40)
        m = _lib.MapIndex(ctx, "ORB", map_des)
        q = _lib.QueryBatch(ctx, "ORB", frames)
        d = torch.zeros((F, I), dtype=torch.int32, device="cuda")
        out = {"config": "ORB %d x %d x 500 x 500" % (F, I)}
        for algo, tag in ((_lib.ALGO_TENSOR, "tensor"), (_lib.ALGO_SIMT, "simt")):
            s0 = ctx.stats()
            r = prof(lambda: m.match_counts_device(q, d.data_ptr(), algo=algo), ("orb", "orb_fixup", "ingest"), reps=3, warm=1)
            s1 = ctx.stats()
            ctx.sync()
            t0 = time.perf_counter()
            for _ in range(3):
                m.match_counts_device(q, d.data_ptr(), algo=algo)
            ctx.sync()
            wall = (time.perf_counter() - t0) / 3
            out[tag] = {"kernel_ms": r["orb"], "fixup_ms": r["orb_fixup"], "prep_ms": r["ingest"], "wall_ms": wall * 1e3,
                        "pairs_per_s": F * I / wall, "candidates_per_call": (s1["orb_candidates"] - s0["orb_candidates"]) / 4,
                        "int8_frac": 2.0 * 512 * 512 * 256 * F * I / (r["orb"] * 1e-3) / i8_peak if tag == "tensor" else None}
            out[tag + "_sum"] = int(d.sum().item())
        print(json.dumps(out), flush=True)
        q.close(); m.close()

if "bow" in names:
    L = 16
    locs = []
    for l in range(L):
        rng = np.random.default_rng(500 + l)
        voc = np.clip(synth.sift_like(rng, 10000) + rng.normal(0, 0.3, size=(10000, 128)), 0, 255).astype(np.float32)
        im = rng.random((100, 10000), dtype=np.float32) * (rng.random((100, 10000)) < 0.03)
        im /= np.maximum(np.linalg.norm(im, axis=1, keepdims=True), 1e-9)
        locs.append((im.astype(np.float32), rng.random(10000, dtype=np.float32) + 0.5, voc))
    bow = _lib.BowIndex(ctx, locs, 10000)
    rng = np.random.default_rng(9)
    slots = ("bow_vq_tc", "bow_resolve", "bow_rescan", "bow_vq", "bow_score", "ingest")
    for F, kind in ((50, "clustered"), (1, "clustered"), (50, "random"), (1, "random")):
        if kind == "clustered":
            frames = [np.clip(np.rint(locs[f % L][2][rng.choice(10000, 300)] + rng.normal(0, 6.0, size=(300, 128))), 0, 255).astype(np.float32)
                      for f in range(F)]
        else:
            frames = [synth.sift_like(rng, 300) for _ in range(F)]
        r = prof(lambda: bow.score(frames), slots, reps=3, warm=1)
        t0 = time.perf_counter()
        for _ in range(3):
            bow.score(frames)
        wall = (time.perf_counter() - t0) / 3
        s_a = ctx.stats()
        bow.score(frames)
        st = ctx.stats()
        st["bow_rescanned_rows"] -= s_a["bow_rescanned_rows"]
        flops = 2.0 * F * 300 * 10000 * 128 * L
        print(json.dumps({"config": "BOW %d frames x %d locations, %s" % (F, L, kind), "e2e_ms": wall * 1e3, **{k + "_ms": v for k, v in r.items()},
                          "tc_fp16_TFLOPs_algorithmic": flops / (r["bow_vq_tc"] * 1e-3) / 1e12, "fp16_peak_TF_assumed": i8_peak / 2e12,
                          "frac_of_fp16_peak_incl_ext": flops * 9 / 8 / (r["bow_vq_tc"] * 1e-3) / (i8_peak / 2),
                          "rescanned_rows_per_call": st["bow_rescanned_rows"], "resolve_entries": st["bow_resolve_entries"], "rows_x_locations": F * 300 * L}), flush=True)
    bow.close()
