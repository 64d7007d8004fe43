"""The reference-facing API on the GPU: pymcl_b200.Matcher / analyzer run the way the reference's README and analyze.py
drive them, compared byte-for-byte with the files the UNMODIFIED reference wrote (tests/golden/) and with the oracle."""
import glob
import json
import os

import numpy as np
import pytest

import oracle
import pipeline
import pymcl_b200  # noqa: F401
from pymcl_b200 import _lib, synth

pytestmark = pytest.mark.gpu
VIEW = (160, 120)


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("mcl_ds"))
    synth.make_dataset(root, seed=0, n_locations=7, n_images=25, n_frames=6, view=VIEW)
    return root


@pytest.fixture()
def in_dataset(dataset, monkeypatch):
    """cwd = dataset root (the reference uses cwd-relative paths) and a sorted glob (the golden run pinned it)."""
    monkeypatch.chdir(dataset)
    orig = glob.glob
    monkeypatch.setattr(glob, "glob", lambda *a, **k: sorted(orig(*a, **k)))
    return dataset


def _same_descriptors_as_golden(golden_dir):
    import cv2
    return json.load(open(os.path.join(golden_dir, "desc_hash.json")))["cv2"] == cv2.__version__


def test_matcher_run_like_the_readme(in_dataset, ctx):
    """README usage: Matcher(alg, width=, height=) + setQuery/setDirectory/setIndex/createFeatureIndex + run()."""
    from pymcl_b200.Matcher import Matcher
    m = Matcher("SIFT", width=VIEW[0], height=VIEW[1])
    m.setDirectory("map/3")
    index = m.createFeatureIndex()
    m.setIndex(index)
    m.setQuery("cam1_img/0002.png")
    total, probs = m.run()
    q = pipeline.query_descriptors("cam1_img/0002.png", "SIFT", *VIEW)
    counts = [oracle.sift_pair_count(q, index["map/3/angle" + str(i).zfill(3) + ".png"][1]) for i in range(0, 375, 15)]
    want_total, want_probs = oracle.run_from_counts(counts)
    assert (total, probs) == (want_total, want_probs)
    assert isinstance(total, int) and all(isinstance(x, float) for x in probs)
    # single-image entry point and the DOR window
    assert m.SIFTMatch("map/3/angle045.png") == counts[3]
    t2, p2 = m.optRun(23)
    win = oracle.opt_window(23)
    want2 = oracle.run_from_counts([c if w else 1 for c, w in zip(counts, win)])
    assert (t2, p2) == want2
    assert m.optRun(None) == (total, probs)


def test_matcher_orb_modes_and_descriptor_entry(in_dataset, ctx):
    from pymcl_b200.Matcher import Matcher
    m = Matcher("ORB", width=VIEW[0], height=VIEW[1])
    m.setDirectory("map/0")
    index = m.createFeatureIndex()
    m.setIndex(index)
    m.setQuery("cam1_img/0000.png")
    q = pipeline.query_descriptors("cam1_img/0000.png", "ORB", *VIEW)
    paths = ["map/0/angle" + str(i).zfill(3) + ".png" for i in range(0, 375, 15)]
    assert m.run() == oracle.run_from_counts([oracle.orb_pair_count_knn(q, index[p][1]) for p in paths])
    m2 = Matcher("ORB", width=VIEW[0], height=VIEW[1])
    m2.orb_mode = "crosscheck"
    m2.setDirectory("map/0")
    m2.setIndex(index)
    m2.setQuery("cam1_img/0000.png")
    assert m2.run() == oracle.run_from_counts([oracle.orb_pair_count_crosscheck(q, index[p][1]) for p in paths])
    # SURF cannot be extracted in this OpenCV build (non-free OFF): descriptor-level entry
    rng = np.random.default_rng(0)
    sidx = dict((p, (None, synth.surf_like(rng, 200))) for p in paths)
    qs = synth.surf_like(rng, 150)
    qs[:60] = synth.surf_noisy_copy(rng, sidx[paths[7]][1][:60])
    m3 = Matcher("SURF", width=640, height=480)
    m3.setDirectory("map/0")
    m3.setIndex(sidx)
    m3.setQueryDescriptors(qs)
    total, probs = m3.run()
    want = [oracle.surf_pair_count(qs, sidx[p][1])[0] for p in paths]
    assert total == sum(want) and int(np.argmax(probs)) == 7


def test_analyzer_sift_files_byte_identical_to_reference(in_dataset, golden_dir, ctx):
    if not _same_descriptors_as_golden(golden_dir):
        pytest.skip("different OpenCV build than the golden run")
    from pymcl_b200.analyze import analyzer
    a = analyzer("SIFT", VIEW[0], VIEW[1])
    a.createRawP()
    assert open("rawP.txt").read() == open(os.path.join(golden_dir, "sift_rawP.txt")).read()
    a.processRaw()
    assert open("out.txt").read() == open(os.path.join(golden_dir, "sift_out.txt")).read()
    assert open("bestGuess.txt").read() == open(os.path.join(golden_dir, "sift_bestGuess.txt")).read()


def test_analyzer_sift_dor_byte_identical_to_reference(in_dataset, golden_dir, ctx):
    if not _same_descriptors_as_golden(golden_dir):
        pytest.skip("different OpenCV build than the golden run")
    from pymcl_b200.analyze import analyzer
    a = analyzer("SIFT", VIEW[0], VIEW[1])
    a.optP()
    assert open("out.txt").read() == open(os.path.join(golden_dir, "sift_dor_out.txt")).read()
    assert open("bestGuess.txt").read() == open(os.path.join(golden_dir, "sift_dor_bestGuess.txt")).read()


@pytest.mark.parametrize("mode,tag", [("knn", "orb_knn"), ("crosscheck", "orb_cc")])
def test_analyzer_orb_rawp_byte_identical_to_reference(in_dataset, golden_dir, ctx, mode, tag):
    if not _same_descriptors_as_golden(golden_dir):
        pytest.skip("different OpenCV build than the golden run")
    from pymcl_b200.analyze import analyzer
    a = analyzer("ORB", VIEW[0], VIEW[1])
    a.orb_mode = mode
    a.createRawP()
    assert open("rawP.txt").read() == open(os.path.join(golden_dir, tag + "_rawP.txt")).read()


def test_analyzer_files_match_oracle_pipeline_whatever_the_opencv_build(in_dataset, ctx, tmp_path):
    """Same comparison against the oracle pipeline run here (does not depend on the golden run's OpenCV version)."""
    from pymcl_b200.analyze import analyzer
    a = analyzer("SIFT", VIEW[0], VIEW[1])
    a.createRawP()
    a.processRaw()
    p = pipeline.create_raw_p(".", "SIFT", *VIEW)
    raw = str(tmp_path / "rawP.txt")
    oracle.write_prob(p, raw)
    assert open("rawP.txt").read() == open(raw).read()
    blur_p, best = pipeline.process_raw(".", raw)
    out = str(tmp_path / "out.txt")
    oracle.write_prob(blur_p, out)
    assert open("out.txt").read() == open(out).read()
    assert a.bestGuess == best


def test_analyzer_public_filter_methods(in_dataset, ctx):
    import copy
    from pymcl_b200.analyze import analyzer
    a = analyzer("SIFT", VIEW[0], VIEW[1])
    rng = np.random.default_rng(1)
    prev = [[float(rng.integers(1, 99)), list(rng.random(25))] for _ in range(7)]
    cur = [[float(rng.integers(1, 99)), list(rng.random(25))] for _ in range(7)]
    assert a.prevWeight(prev, cur) == oracle.prev_weight(prev, cur)
    for blur in (12.5, 200.0, 950.0):
        assert a.probUpdate(prev, cur, blur) == oracle.prob_update(prev, cur, blur)
    for cmd in "lrfbs":
        p1, p2 = copy.deepcopy(prev), copy.deepcopy(prev)
        assert a.accountCommand(cmd, p1) == oracle.account_command(cmd, p2)
        assert p1 == p2  # same in-place mutation as the reference's shallow copy
    import cv2
    v = a.Laplacian("cam1_img/0001.png")
    assert v == oracle.laplacian_var(cv2.imread("cam1_img/0001.png", 0)) and isinstance(v, float)


def test_bow_index_and_match_through_the_matcher(in_dataset, ctx):
    """createIndex (scipy k-means + GPU word assignment) -> .pkl in the reference's format -> BOWMatch / run()."""
    import joblib
    from pymcl_b200.Matcher import Matcher
    np.random.seed(5)
    m = Matcher("BOW", width=VIEW[0], height=VIEW[1])
    m.numWords = 150
    m.createIndex("map/1")
    im_features, image_paths, idf, num_words, voc = joblib.load("map/1.pkl")
    assert im_features.shape == (25, 150) and idf.shape == (150,) and voc.shape[1] == 128 and num_words == 150
    m.setDirectory("map/1")
    m.setQuery("cam1_img/0002.png")
    score = m.BOWMatch("map/1.pkl")
    q = pipeline.query_descriptors("cam1_img/0002.png", "SIFT", *VIEW)
    words_ref, score_ref = oracle.scipy_bow_score(q, voc, idf, im_features, 150)
    assert score.shape == (1, 25) and score.dtype == np.float32
    # words may differ from scipy's only on float32 near-ties; the score is checked unconditionally against the reference's
    # arithmetic on the words the GPU chose, and against scipy's own score when the words agree
    ow, gap = oracle.bow_words(q, voc)
    bow = _lib.BowIndex(ctx, [(im_features, idf, voc)], 150)
    s2, words = bow.score([q], want_words=True)
    bow.close()
    assert np.array_equal(s2, score)
    assert np.all(gap[words[0] != ow] < 1e-5) and np.all(gap[words[0] != words_ref] < 1e-5)
    assert np.allclose(score, oracle.bow_score_from_words(words[0], idf, im_features, 150), rtol=0, atol=1e-6)
    if np.array_equal(words[0], words_ref):
        assert np.allclose(score, score_ref, rtol=0, atol=1e-6)
    total, probs = m.run()
    assert total == 10 * np.max(score) and probs == score[0].tolist()


def test_matcher_color_search_and_run_equal_the_reference(in_dataset, golden_dir, ctx):
    """Matcher('Color'): createColorIndex / setColorIndex / colorSearch / run, against the reference's own values."""
    if not _same_descriptors_as_golden(golden_dir):
        pytest.skip("different OpenCV build than the golden run")
    from pymcl_b200.Matcher import Matcher
    g = np.load(os.path.join(golden_dir, "color_case.npz"))
    m = Matcher("Color", width=VIEW[0], height=VIEW[1])
    m.setDirectory("map/0")
    index = m.createColorIndex()
    m.setColorIndex(index)
    assert sorted(index.keys()) == ["angle" + str(i).zfill(3) + ".png" for i in range(0, 375, 15)]
    for f, frame in enumerate(sorted(glob.glob("cam1_img/*.png"))):
        m.setQuery(frame)
        assert np.array_equal(m.createHistogram(m.image), g["q"][f])
        results = m.colorSearch()
        assert [r[1] for r in results] == [n for _, n in sorted(zip(g["chi"][f, :25].tolist(), sorted(index.keys())))]
        assert all(isinstance(r[0], np.float32) for r in results)
        total, probs = m.run()
        assert isinstance(total, np.float32) and total.tobytes() == g["run_total"][f].tobytes()
        assert np.array_equal(np.asarray(probs, dtype=np.float32).view(np.uint32), g["run_probs"][f].view(np.uint32))
        assert m.optRun(7) == (total, probs)      # the colour branch ignores the angle window (Matcher.py:377-385)


def test_analyzer_color_files_byte_identical_to_reference(in_dataset, golden_dir, ctx):
    if not _same_descriptors_as_golden(golden_dir):
        pytest.skip("different OpenCV build than the golden run")
    from pymcl_b200.analyze import analyzer
    a = analyzer("Color", VIEW[0], VIEW[1])
    a.createRawP()
    assert open("rawP.txt").read() == open(os.path.join(golden_dir, "color_rawP.txt")).read()
    a.processRaw()
    assert open("out.txt").read() == open(os.path.join(golden_dir, "color_out.txt")).read()
    assert open("bestGuess.txt").read() == open(os.path.join(golden_dir, "color_bestGuess.txt")).read()


def test_analyzer_color_dor_byte_identical_to_reference(in_dataset, golden_dir, ctx):
    """optP with the colour method: np.float32 results flow through the filter (mcl_filter_step_typed)."""
    if not _same_descriptors_as_golden(golden_dir):
        pytest.skip("different OpenCV build than the golden run")
    from pymcl_b200.analyze import analyzer
    a = analyzer("Color", VIEW[0], VIEW[1])
    a.optP()
    assert open("out.txt").read() == open(os.path.join(golden_dir, "color_dor_out.txt")).read()
    assert open("bestGuess.txt").read() == open(os.path.join(golden_dir, "color_dor_bestGuess.txt")).read()


def test_analyzer_color_matches_oracle_pipeline_whatever_the_opencv_build(in_dataset, ctx, tmp_path):
    from pymcl_b200.analyze import analyzer
    a = analyzer("Color", VIEW[0], VIEW[1])
    a.createRawP()
    raw = str(tmp_path / "rawP.txt")
    oracle.write_prob(pipeline.create_raw_p_color(in_dataset, *VIEW), raw)
    assert open("rawP.txt").read() == open(raw).read()
    a = analyzer("Color", VIEW[0], VIEW[1])
    a.optP()
    blur_p, best = pipeline.opt_p_color(in_dataset, *VIEW)
    out = str(tmp_path / "out.txt")
    oracle.write_prob(blur_p, out)
    assert open("out.txt").read() == open(out).read()
    assert a.bestGuess == best


def test_analyzer_sidecar_gives_the_same_files(in_dataset, golden_dir, ctx):
    """processRaw fed from rawP.npz instead of the text parse writes the same out.txt / bestGuess.txt."""
    from pymcl_b200.analyze import analyzer
    a = analyzer("SIFT", VIEW[0], VIEW[1])
    a.sidecar = True
    a.createRawP()
    z = np.load("rawP.npz")
    assert z["counts"].shape == (6, 175) and z["counts"].dtype == np.int32
    assert a.readSidecar("rawP.txt") == a.readProb("rawP.txt")
    a.processRaw()
    with_sidecar = (open("out.txt").read(), open("bestGuess.txt").read())
    b = analyzer("SIFT", VIEW[0], VIEW[1])
    b.processRaw()
    assert (open("out.txt").read(), open("bestGuess.txt").read()) == with_sidecar
    if _same_descriptors_as_golden(golden_dir):
        assert with_sidecar[0] == open(os.path.join(golden_dir, "sift_out.txt")).read()
    os.remove("rawP.npz")


def test_analyzer_sift_800x600_rawp_byte_identical_to_reference(tmp_path_factory, monkeypatch, golden_dir, ctx):
    """The README's image size (README.md:44; SURVEY 8d config 3, image-level variant): 7 x 25 textured 800x600 map views with
    1-3 k SIFT descriptors each + 4 frames.  tests/golden/sift800_* was written by the UNMODIFIED reference (exact search,
    oracle/make_golden.py --big); the images are regenerated from the seed and re-extracted here."""
    import cv2
    g = np.load(os.path.join(golden_dir, "sift800_case.npz"))
    if str(g["cv2"]) != cv2.__version__:
        pytest.skip("different OpenCV build than the golden run")
    from pymcl_b200.analyze import analyzer
    from pymcl_b200.Matcher import Matcher
    root = str(tmp_path_factory.mktemp("mcl_ds800"))
    synth.make_dataset(root, seed=8, n_locations=7, n_images=25, n_frames=4, view=(800, 600), n_shapes=5600)
    monkeypatch.chdir(root)
    orig = glob.glob
    monkeypatch.setattr(glob, "glob", lambda *a, **k: sorted(orig(*a, **k)))
    a = analyzer("SIFT", 800, 600)
    a.createRawP()
    n_desc = [len(a.indices[loc][p][1]) for loc in range(7) for p in sorted(a.indices[loc].keys())]
    assert np.array_equal(n_desc, g["n_desc"]) and min(n_desc) >= 1000
    assert open("rawP.txt").read() == open(os.path.join(golden_dir, "sift800_rawP.txt")).read()
    # the reference's own SIFTMatch counts of location 0, through Matcher.SIFTMatch
    m = Matcher("SIFT", width=800, height=600)
    m.setDirectory("map/0")
    m.setIndex(a.indices[0])
    paths = ["map/0/angle" + str(i).zfill(3) + ".png" for i in range(0, 375, 15)]
    for f, frame in enumerate(sorted(orig("cam1_img/*.png"))):
        m.setQuery(frame)
        assert [m.SIFTMatch(p) for p in paths] == g["counts"][f].tolist()


def test_multi_gpu_nccl_drivers_and_sharded_paths(tmp_path):
    """One process per GPU over NCCL (scripts/multi_gpu_check.py under torchrun): analyzer SIFT / colour / BOW drivers against
    the reference's golden files, and the sharded descriptor-level paths (rows uploaded once per box + NVLink all-gather,
    device all-gather of the counts) against the single-GPU result.  Skips on a box with one GPU."""
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env["MCL_MULTI_GPU_ROOT"] = str(tmp_path)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(repo, "scripts", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    log = r.stdout + "\n--- stderr ---\n" + r.stderr[-4000:]
    out_dir = os.path.join(repo, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    open(os.path.join(out_dir, "multi_gpu_check.log"), "w").write(log)
    assert r.returncode == 0, log
    for tag in ("MULTI_GPU_OK", "MULTI_GPU_COLOR_OK", "MULTI_GPU_BOW_OK", "MULTI_GPU_SHARDED_UPLOAD_OK", "MULTI_GPU_ORB_OK"):
        assert tag in r.stdout, log
