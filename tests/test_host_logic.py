"""Host-side logic that needs no GPU: packing, sharding, the DOR window, the reference-facing signatures, and the
world_size-2 (gloo) path of ShardedMap with an oracle-backed local matcher standing in for the GPU."""
import inspect
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
import pymcl_b200  # noqa: F401
from pymcl_b200 import _lib, engine, synth
from pymcl_b200.Matcher import Matcher, angle_window
from conftest import REPO


def test_angle_window_equals_reference_logic():
    for k in range(25):
        assert angle_window(k) == oracle.opt_window(k)


def test_matcher_signature_and_positional_quirk():
    """Matcher(algorithm, index=None, width=800, height=600): positional ('SIFT', 320, 240) sets index=320, w=240."""
    sig = inspect.signature(Matcher.__init__)
    positional = [n for n, q in sig.parameters.items() if q.kind == q.POSITIONAL_OR_KEYWORD]
    assert positional == ["self", "algorithm", "index", "width", "height"]      # the reference's signature, Matcher.py:41
    keyword_only = {n: q.default for n, q in sig.parameters.items() if q.kind == q.KEYWORD_ONLY}
    assert keyword_only == {"imagesPerLocation": 25, "angleStep": 15}           # the reference's constants as defaults
    m = Matcher("SIFT", 320, 240)
    assert (m.index, m.w, m.h, m.numWords) == (320, 240, 600, 10000)
    m = Matcher("ORB", width=320, height=240)
    assert (m.index, m.w, m.h) == (None, 320, 240)
    for name in ("setQuery", "setDirectory", "setIndex", "createFeatureIndex", "run", "optRun", "SIFTMatch", "SURFMatch",
                 "ORBMatch", "BOWMatch", "createIndex", "writeIndices"):
        assert callable(getattr(Matcher, name))


def test_analyzer_signature():
    from pymcl_b200.analyze import analyzer
    sig = inspect.signature(analyzer.__init__)
    assert list(sig.parameters)[:5] == ["self", "method", "width", "height", "numLocations"]
    assert sig.parameters["numLocations"].default == 7 and sig.parameters["imagesPerLocation"].default == 25
    assert sig.parameters["angleStep"].default == 15 and sig.parameters["imagesPerLocation"].kind == inspect.Parameter.KEYWORD_ONLY
    for name in ("createIndex", "createRawP", "processRaw", "optP", "probUpdate", "prevWeight", "accountCommand",
                 "writeProb", "readProb", "readCommand", "readBestGuess", "Laplacian", "writeSidecar", "readSidecar"):
        assert callable(getattr(analyzer, name))


def test_matcher_color_host_logic_equals_oracle(golden_dir):
    """The colour branch of run() (float32 scalar arithmetic in ascending-distance order) against the oracle restatement
    and the reference's own numbers; createHistogram / setColorIndex / createColorIndex exist with the reference's names."""
    g = np.load(os.path.join(golden_dir, "color_case.npz"))
    names = ["angle" + str(i).zfill(3) + ".png" for i in range(0, 375, 15)]
    m = Matcher("Color", width=160, height=120)
    m.setDirectory("map/0")
    for f in range(len(g["q"])):
        results = sorted([(v, k) for k, v in zip(names, g["chi"][f, :25])])
        total, probs = m._color_finish(results)
        want_total, want_probs = oracle.color_run(results, "map/0")
        assert type(total) is np.float32 and total.tobytes() == np.float32(want_total).tobytes()
        assert [p.tobytes() for p in probs] == [p.tobytes() for p in want_probs]
        assert total.tobytes() == g["run_total"][f].tobytes()
    for name in ("createHistogram", "createColorIndex", "setColorIndex", "colorSearch"):
        assert callable(getattr(Matcher, name))


def test_sidecar_round_trips_what_readprob_parses(tmp_path, monkeypatch):
    """rawP.npz: the dictionary it yields equals readProb() of the text it mirrors -- for integer counts,
    float32 values (colour / BOW) and Python floats -- and it is ignored once the text changes."""
    from pymcl_b200.analyze import analyzer
    monkeypatch.chdir(tmp_path)
    open("commands.txt", "w").write("0000 s\n")
    a = analyzer("SIFT", 160, 120, numLocations=3)
    rng = np.random.default_rng(4)
    p = []
    for f in range(4):
        counts = rng.integers(0, 50, size=(3, 25))
        for l in range(3):
            if f == 0:
                p.append(a._run_result(counts[l]))
            elif f == 1:
                v = rng.random(25).astype(np.float32)
                p.append([np.float32(10) * v.max(), list(v)])
            elif f == 2:
                p.append([1, [1 / 75] * 25])
            else:
                p.append([float(rng.random() * 1e-7), list(rng.random(25) * 1e5)])
    a.writeProb(p, "rawP.txt", "w")
    assert a.readSidecar("rawP.txt") is None
    a.writeSidecar(p, "rawP.txt", counts=np.zeros((4, 75), np.int32))
    assert a.readSidecar("rawP.txt") == a.readProb("rawP.txt")
    assert np.load("rawP.npz")["counts"].shape == (4, 75)
    open("rawP.txt", "a").write("1\n[1.0]\n")
    assert a.readSidecar("rawP.txt") is None
    # bag of words: every location owns its own number of images, so the probability lists are ragged
    ragged = []
    for f in range(2):
        for n in (4, 9, 6):
            v = rng.random(n).astype(np.float32)
            ragged.append([np.float32(10) * v.max(), list(v)])
    a.writeProb(ragged, "rawP.txt", "w")
    a.writeSidecar(ragged, "rawP.txt")
    got = a.readSidecar("rawP.txt")
    assert [len(c[1]) for c in got["0001"]] == [4, 9, 6]
    # (readProb's own parse of the same text, for the values)
    assert got == a.readProb("rawP.txt")


def test_large_batches_are_staged_without_changing_the_bytes():
    """pack_descriptors concatenates big batches into a reusable staging buffer (pinned when a CUDA runtime is there), on a
    few threads: the result must equal np.concatenate, for mixed dtypes too, and stay valid until the next call."""
    rng = np.random.default_rng(3)
    frames = [rng.integers(0, 256, size=(int(n), 128)).astype(np.float32) for n in rng.integers(1500, 2500, size=24)]
    frames[5] = None
    frames[7] = frames[7].astype(np.uint8)          # mixed dtypes are converted to float32
    off, cat, code = _lib.pack_descriptors(_lib.SIFT, frames)
    real = [f for f in frames if f is not None]
    assert code == _lib.F32 and cat.dtype == np.float32 and cat.flags.c_contiguous
    assert cat.nbytes >= (8 << 20)                   # the threaded path
    assert np.array_equal(cat, np.concatenate([f.astype(np.float32) for f in real]))
    assert list(off) == list(np.cumsum([0] + [0 if f is None else len(f) for f in frames]))
    first = cat.copy()
    off2, cat2, _ = _lib.pack_descriptors(_lib.SIFT, frames[:12])   # reuses the buffer: only the latest result is valid
    assert np.array_equal(cat2, first[:off2[-1]])
    small = _lib.pack_descriptors(_lib.ORB, [np.zeros((3, 32), np.uint8), np.ones((2, 32), np.uint8)])[1]
    assert small.shape == (5, 32) and small[3:].min() == 1


def test_typed_lists_round_trip_through_the_filter_packing():
    """analyzer._pack/_unpack: the scalar-type tags (Python number / np.float64 / np.float32) survive the trip to the
    (L, 1+I) float64 matrix the typed filter kernel works on, with exact values."""
    from pymcl_b200.analyze import analyzer
    lists = [[1, [1 / 75] * 25], [np.float32(21.411722), [np.float32(0.1)] * 25], [np.float64(3.5), [np.float64(0.25)] * 25],
             [np.float32(7.0), [0.5] * 25]]
    values, tags = analyzer._pack(lists)
    assert values.shape == (4, 26) and tags.tolist() == [[0, 0], [2, 2], [1, 1], [2, 0]]
    back = analyzer._unpack(values, tags)
    assert type(back[0][0]) is float and type(back[0][1][0]) is float and back[0][1][0] == 1 / 75
    assert type(back[1][0]) is np.float32 and back[1][0] == np.float32(21.411722) and type(back[1][1][3]) is np.float32
    assert type(back[2][0]) is np.float64 and type(back[3][0]) is np.float32 and type(back[3][1][0]) is float


def test_pack_descriptors():
    rng = np.random.default_rng(0)
    a, b = synth.sift_like(rng, 5), synth.sift_like(rng, 3)
    off, cat, code = _lib.pack_descriptors(_lib.SIFT, [a, None, b, np.zeros((0, 128), np.float32)])
    assert off.tolist() == [0, 5, 5, 8, 8] and cat.shape == (8, 128) and code == _lib.F32 and cat.dtype == np.float32
    off, cat, code = _lib.pack_descriptors(_lib.SIFT, [a.astype(np.uint8)])
    assert code == _lib.U8 and cat.dtype == np.uint8
    off, cat, code = _lib.pack_descriptors(_lib.ORB, [None, None])
    assert off.tolist() == [0, 0, 0] and cat.shape == (0, 32)
    with pytest.raises(_lib.MclError):
        _lib.pack_descriptors(_lib.ORB, [np.zeros((4, 31), np.uint8)])


def test_shard_bounds_cover_and_balance():
    rng = np.random.default_rng(1)
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 5, 64, 175, 3200):
            rows = rng.integers(0, 3000, size=n).tolist()
            b = engine.shard_bounds(rows, world)
            assert len(b) == world and b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            if n >= 64 * world:
                w = [sum(rows[l:h]) + (h - l) for l, h in b]
                assert max(w) <= 1.25 * (sum(w) / world) + 3001
    assert engine.shard_bounds([2048] * 3200, 8) == [(i * 400, (i + 1) * 400) for i in range(8)]


class OracleLocal(object):
    """Stands in for _lib.MapIndex in CPU tests: same match_counts contract, computed by the oracle."""

    def __init__(self, descs):
        self.descs = descs
        self.n_images = len(descs)

    def match_counts(self, q_descs, mode=0, ratio=0.7, mask=None, fill_value=1, algo=0, packed=None):
        if packed is not None:   # (offsets, matrix, dtype code): the rows of all frames in one buffer
            off, cat, _ = packed
            q_descs = [cat[off[f]:off[f + 1]] for f in range(len(off) - 1)]
        out = np.zeros((len(q_descs), len(self.descs)), dtype=np.int32)
        for f, q in enumerate(q_descs):
            for i, m in enumerate(self.descs):
                if mask is not None and not mask[f, i]:
                    out[f, i] = fill_value
                else:
                    out[f, i] = oracle.sift_pair_count(q, m, ratio)
        return out


class OracleBow:
    """Stands in for _lib.BowIndex on CPU: BOWMatch of every query against every location through the oracle."""

    def __init__(self, locations, num_words):
        self.locations = locations
        self.W = num_words

    def score(self, q_descs, want_words=False, packed=None):
        if packed is not None:
            off, cat, _ = packed
            q_descs = [cat[off[f]:off[f + 1]] for f in range(len(off) - 1)]
        cols = sum(len(l[0]) for l in self.locations)
        out = np.zeros((len(q_descs), cols), dtype=np.float32)
        for f, q in enumerate(q_descs):
            c = 0
            for imf, idf, voc in self.locations:
                if q is not None and len(q):
                    words, _ = oracle.bow_words(np.asarray(q, np.float32), voc)
                    out[f, c:c + len(imf)] = oracle.bow_score_from_words(words, idf, imf, self.W)[0]
                c += len(imf)
        return out, None


def synthetic_bow_locations(seed, n_loc, W=40):
    rng = np.random.default_rng(seed)
    locs = []
    for i in range(n_loc):
        voc = synth.sift_like(rng, W - (i % 3)).astype(np.float32) + rng.random((W - (i % 3), 128)).astype(np.float32)
        n_img = 3 + (i * 7) % 5
        imf = rng.random((n_img, W)).astype(np.float32) * (rng.random((n_img, W)) < 0.3)
        imf /= np.maximum(np.linalg.norm(imf, axis=1, keepdims=True), 1e-9)
        idf = np.log((n_img + 1.0) / (rng.integers(0, n_img + 1, W) + 1.0)).astype(np.float32)
        locs.append((imf.astype(np.float32), idf, voc))
    return locs


def test_sharded_bow_single_process_matches_unsharded():
    locs = synthetic_bow_locations(1, 5)
    rng = np.random.default_rng(2)
    frames = [synth.sift_like(rng, 30), None, synth.sift_like(rng, 11)]
    sb = engine.ShardedBow(len(locs), lambda i: locs[i], local_factory=OracleBow)
    want = OracleBow(locs, 40).score(frames)[0]
    got = sb.score(frames)
    assert got.shape == (3, sum(len(l[0]) for l in locs)) and np.array_equal(got, want)
    assert list(sb.img_off) == list(np.cumsum([0] + [len(l[0]) for l in locs]))


def test_sharded_map_single_process_matches_unsharded():
    map_des, frames = synth.sift_map_and_frames(2, 9, 60, 3, 50)
    sm = engine.ShardedMap("SIFT", map_des, local_factory=OracleLocal)
    want = OracleLocal(map_des).match_counts(frames)
    assert np.array_equal(sm.match_counts(frames), want)
    # the same batch handed over as one packed buffer (what bench.py's end-to-end path does with pinned memory)
    packed = _lib.pack_descriptors(_lib.SIFT, frames)
    assert np.array_equal(sm.match_counts(None, packed=packed), want)


def test_sharded_bow_accepts_a_packed_batch():
    locs = synthetic_bow_locations(4, 3)
    rng = np.random.default_rng(5)
    frames = [synth.sift_like(rng, 17), synth.sift_like(rng, 9), None]
    sb = engine.ShardedBow(len(locs), lambda i: locs[i], local_factory=OracleBow)
    want = sb.score(frames)
    packed = _lib.pack_descriptors(_lib.BOWQ, frames)
    assert packed[0].tolist() == [0, 17, 26, 26]
    assert np.array_equal(sb.score(None, packed=packed), want)


WORKER = r'''
import os, sys
sys.path.insert(0, {repo!r}); sys.path.insert(0, os.path.join({repo!r}, "oracle")); sys.path.insert(0, os.path.join({repo!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
import pymcl_b200
from pymcl_b200 import engine, synth
from test_host_logic import OracleLocal, OracleBow, synthetic_bow_locations
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
sizes = [40, 0, 75, 33, 64, 1, 90, 55, 20, 61, 48]
rng = np.random.default_rng(3)
map_des = [synth.sift_like(rng, n) if n else None for n in sizes]
frames = [synth.sift_like(rng, 70), synth.sift_like(rng, 30), None]
frames[0][:30] = synth.sift_noisy_copy(rng, map_des[6][:30], 4.0)
frames[1][:20] = synth.sift_noisy_copy(rng, map_des[2][:20], 4.0)
sm = engine.ShardedMap("SIFT", map_des, local_factory=OracleLocal)
assert sm.world == world and sm.bounds[rank] == (sm.lo, sm.hi)
mask = (np.random.default_rng(4).random((3, len(sizes))) < 0.6).astype(np.uint8)
got = sm.match_counts(frames)
got_m = sm.match_counts(frames, mask=mask, fill_value=1)
want = OracleLocal(map_des).match_counts(frames)
assert np.array_equal(got, want), (rank, got, want)
assert np.array_equal(got_m, np.where(mask > 0, want, 1))
assert want[0, 6] >= 20 and want[1, 2] >= 10
# every rank holds a strict subset and only the counts travelled
assert sm.hi - sm.lo < len(sizes)
# extraction-style work dealt round-robin to the ranks, results gathered in order on every rank
items = list(range(11))
res = engine.sharded_apply(lambda i: (np.full((i % 3, 4), i, np.uint8) if i % 4 else None), items)
assert len(res) == 11 and res[4] is None and res[5].shape == (2, 4) and int(res[10][0, 0]) == 10
# bag of words: LOCATIONS are sharded (each owns its code book + index), float32 scores all-gathered
locs = synthetic_bow_locations(1, 5)
loaded = []
sb = engine.ShardedBow(len(locs), lambda i: (loaded.append(i), locs[i])[1], local_factory=OracleBow)
assert loaded == list(range(*sb.loc_bounds[rank])) and 0 < len(loaded) < len(locs)   # only this rank's indices were loaded
bf = [synth.sift_like(rng, 30), None, synth.sift_like(rng, 11)]
got_b = sb.score(bf)
want_b = OracleBow(locs, 40).score(bf)[0]
assert got_b.dtype == np.float32 and np.array_equal(got_b, want_b), rank
assert list(sb.img_off) == list(np.cumsum([0] + [len(l[0]) for l in locs]))
dist.barrier()
if rank == 0:
    print("SHARD_OK", sm.bounds, sb.loc_bounds)
dist.destroy_process_group()
'''


def test_sharded_map_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(repo=REPO))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "SHARD_OK" in r.stdout


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the reference's CPU routine on the host cores) needs no GPU: one JSON line with the
    contract's keys, impl = reference, a cpu_baseline describing the run and an e2e block with zero transfer bytes."""
    import json
    cmd = [sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
           "--frames", "1", "--images", "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "map-image matches/sec" and line["unit"] == "pairs/s"
    assert line["value"] > 0 and line["cpu_baseline"]["value"] == line["value"] and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and "workload" in line["config"]
    assert line["e2e"] == {"value": line["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


class FakeContext:
    """Stands in for _lib.Context on CPU: every GPU entry point the colour drivers use, answered by the oracle.  Lets the
    CPU suite run analyzer('Color') end to end (host logic, text formats, typed hand-over to the filter) against the files
    the unmodified reference wrote."""

    def chi2(self, q, m, eps=1e-10):
        return oracle.chi2_matrix(q, m, eps)

    def laplacian_var(self, imgs):
        return np.array([oracle.laplacian_var(i) for i in imgs], dtype=np.float64)

    @staticmethod
    def _lists(values, tags):
        conv = {_lib.T_PY: float, _lib.T_F64: np.float64, _lib.T_F32: np.float32}
        return [[conv[int(t0)](row[0]), [conv[int(t1)](v) for v in row[1:]]] for row, (t0, t1) in zip(values, tags)]

    def filter_step(self, stages, cmd, blur, prev, cur, angle_step=15):
        assert stages == _lib.F_ALL
        z = np.zeros((len(prev), 2), np.int8)
        p2, out, _, best = self.filter_step_typed(stages, cmd, blur, _lib.T_PY, prev, z, cur, z, angle_step=angle_step)
        return p2, out, best

    def filter_step_typed(self, stages, cmd, blur, blur_type, prev, prev_type, cur, cur_type, angle_step=15):
        assert stages == _lib.F_ALL
        p, c = self._lists(prev, prev_type), self._lists(cur, cur_type)
        adj, best = oracle.filter_step(cmd, p, c, np.float64(blur) if blur_type == _lib.T_F64 else float(blur), angle_step)
        tag = lambda x: _lib.T_F32 if isinstance(x, np.float32) else (_lib.T_F64 if isinstance(x, np.float64) else _lib.T_PY)
        pack = lambda rows: np.array([[r[0]] + list(r[1]) for r in rows], dtype=np.float64)
        return pack(p), pack(adj), np.array([[tag(r[0]), tag(r[1][0])] for r in adj], np.int8), tuple(best)


def test_analyzer_color_drivers_on_a_fake_context(tmp_path, monkeypatch, golden_dir):
    import glob
    import json
    import cv2
    if json.load(open(os.path.join(golden_dir, "desc_hash.json")))["cv2"] != cv2.__version__:
        pytest.skip("different OpenCV build than the golden run")
    from pymcl_b200.analyze import analyzer
    root = str(tmp_path / "ds")
    synth.make_dataset(root, seed=0, n_locations=7, n_images=25, n_frames=6, view=(160, 120))
    monkeypatch.chdir(root)
    orig = glob.glob
    monkeypatch.setattr(glob, "glob", lambda *a, **k: sorted(orig(*a, **k)))
    monkeypatch.setattr(_lib, "default_context", lambda: FakeContext())
    read = lambda name: open(name).read()
    gold = lambda name: open(os.path.join(golden_dir, name)).read()
    a = analyzer("Color", 160, 120)
    a.createRawP()
    assert read("rawP.txt") == gold("color_rawP.txt")
    a.processRaw()
    assert read("out.txt") == gold("color_out.txt") and read("bestGuess.txt") == gold("color_bestGuess.txt")
    b = analyzer("Color", 160, 120)
    b.optP()
    assert read("out.txt") == gold("color_dor_out.txt") and read("bestGuess.txt") == gold("color_dor_bestGuess.txt")


def test_analyzer_sift_drivers_on_a_fake_context(tmp_path, monkeypatch, golden_dir):
    """analyzer('SIFT') end to end on CPU: cv2 extraction on the thread pool, batched createRawP, the DOR masks of optP,
    the particle filter hand-over and the text writers -- with the GPU entry points answered by the oracle -- against the
    files the unmodified reference wrote."""
    import glob
    import json
    import cv2
    if json.load(open(os.path.join(golden_dir, "desc_hash.json")))["cv2"] != cv2.__version__:
        pytest.skip("different OpenCV build than the golden run")
    from pymcl_b200.analyze import analyzer
    root = str(tmp_path / "ds")
    synth.make_dataset(root, seed=0, n_locations=7, n_images=25, n_frames=6, view=(160, 120))
    monkeypatch.chdir(root)
    orig = glob.glob
    monkeypatch.setattr(glob, "glob", lambda *a, **k: sorted(orig(*a, **k)))
    monkeypatch.setattr(_lib, "default_context", lambda: FakeContext())
    monkeypatch.setattr(_lib, "MapIndex", lambda ctx, kind, descs: OracleLocal(descs))
    read = lambda name: open(name).read()
    gold = lambda name: open(os.path.join(golden_dir, name)).read()
    a = analyzer("SIFT", 160, 120)
    a.sidecar = True
    a.createRawP()
    assert read("rawP.txt") == gold("sift_rawP.txt")
    a.processRaw()          # through the binary side-car
    assert read("out.txt") == gold("sift_out.txt") and read("bestGuess.txt") == gold("sift_bestGuess.txt")
    b = analyzer("SIFT", 160, 120)
    b.optP()
    assert read("out.txt") == gold("sift_dor_out.txt") and read("bestGuess.txt") == gold("sift_dor_bestGuess.txt")


class OracleMap(object):
    """_lib.MapIndex(ctx, kind, descs) answered by the oracle, any kind and mode."""

    def __init__(self, ctx, kind, descs):
        self.kind, self.descs, self.n_images = kind, descs, len(descs)

    def match_counts(self, q_descs, mode=0, ratio=0.7, mask=None, fill_value=1, algo=0):
        if self.kind == "ORB":
            pair = oracle.orb_pair_count_crosscheck if mode == _lib.CROSSCHECK else (lambda q, m: oracle.orb_pair_count_knn(q, m, ratio))
        else:
            pair = lambda q, m: oracle.sift_pair_count(q, m, ratio)
        out = np.zeros((len(q_descs), len(self.descs)), dtype=np.int32)
        for f, q in enumerate(q_descs):
            for i, m in enumerate(self.descs):
                out[f, i] = fill_value if (mask is not None and not mask[f, i]) else pair(q, m)
        return out


@pytest.mark.parametrize("mode,tag", [("knn", "orb_knn"), ("crosscheck", "orb_cc")])
def test_analyzer_orb_driver_on_a_fake_context(tmp_path, monkeypatch, golden_dir, mode, tag):
    import glob
    import json
    import cv2
    if json.load(open(os.path.join(golden_dir, "desc_hash.json")))["cv2"] != cv2.__version__:
        pytest.skip("different OpenCV build than the golden run")
    from pymcl_b200.analyze import analyzer
    root = str(tmp_path / "ds")
    synth.make_dataset(root, seed=0, n_locations=7, n_images=25, n_frames=6, view=(160, 120))
    monkeypatch.chdir(root)
    orig = glob.glob
    monkeypatch.setattr(glob, "glob", lambda *a, **k: sorted(orig(*a, **k)))
    monkeypatch.setattr(_lib, "default_context", lambda: FakeContext())
    monkeypatch.setattr(_lib, "MapIndex", OracleMap)
    a = analyzer("ORB", 160, 120)
    a.orb_mode = mode
    a.createRawP()
    assert open("rawP.txt").read() == open(os.path.join(golden_dir, tag + "_rawP.txt")).read()


def test_analyzer_generic_map_shape_on_a_fake_context(tmp_path, monkeypatch):
    """Keyword-only generalisations of the reference's hard-coded 7 locations x 25 images x 15 degrees (analyze.py:35,107,
    Matcher.py:309): 3 locations x 10 images x 36 degrees, frames matched in chunks of 3 -- createRawP / processRaw / optP
    against the oracle pipeline with the same shape (BASELINE configs 2, 3 and 5 are such generalisations)."""
    import glob
    import pipeline
    from pymcl_b200.analyze import analyzer
    root = str(tmp_path / "ds")
    synth.make_dataset(root, seed=3, n_locations=3, n_images=10, n_frames=5, view=(160, 120), angle_step=36)
    monkeypatch.chdir(root)
    orig = glob.glob
    monkeypatch.setattr(glob, "glob", lambda *a, **k: sorted(orig(*a, **k)))
    monkeypatch.setattr(_lib, "default_context", lambda: FakeContext())
    monkeypatch.setattr(_lib, "MapIndex", OracleMap)
    shape = dict(n_locations=3, n_images=10, step=36)
    a = analyzer("ORB", 160, 120, numLocations=3, imagesPerLocation=10, angleStep=36, framesPerCall=3)
    a.createRawP()
    oracle.write_prob(pipeline.create_raw_p(root, "ORB", 160, 120, **shape), "want_rawP.txt")
    assert open("rawP.txt").read() == open("want_rawP.txt").read()
    assert len(a.rawP) == 5 * 3 and all(len(c[1]) == 10 for c in a.rawP)
    a.processRaw()
    blur_p, best = pipeline.process_raw(root, "rawP.txt", **shape)
    oracle.write_prob(blur_p, "want_out.txt")
    assert open("out.txt").read() == open("want_out.txt").read() and a.bestGuess == best
    b = analyzer("ORB", 160, 120, numLocations=3, imagesPerLocation=10, angleStep=36)
    b.optP()
    blur_p, best = pipeline.opt_p(root, "ORB", 160, 120, **shape)
    oracle.write_prob(blur_p, "want_dor.txt")
    assert open("out.txt").read() == open("want_dor.txt").read() and b.bestGuess == best
    m = Matcher("ORB", width=160, height=120, imagesPerLocation=10, angleStep=36)
    m.setDirectory("map/1")
    assert m._angle_paths() == ["map/1/angle%03d.png" % (36 * k) for k in range(10)]
