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
    ow, gap = oracle.bow_words(q, voc)
    if np.all(gap > 1e-5):
        assert np.allclose(score, score_ref, rtol=0, atol=2e-6)
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
