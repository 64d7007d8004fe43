'''
General-Purpose Image Matching on B200
======================================

Drop-in for the reference's ``Matcher`` (reference Matcher.py:35-388): same constructor, same
``setQuery/setDirectory/setIndex/createFeatureIndex/run/optRun/*Match/BOWMatch/createIndex/writeIndices``
surface, same return values -- but the brute-force descriptor matching behind ``run()`` (2-nearest
neighbours, Lowe's ratio test, per-image good-match count) runs in hand-written sm_100a kernels
(libmcl.so, include/mcl.h) on descriptors that stay resident in HBM.

What is deliberately the same as the reference
  * keypoints / descriptors come from the reference's cv2 extractors (inputs are identical);
  * ``setQuery`` = imread -> bilateralFilter(9,75,75) -> resize((w,h))      (Matcher.py:48-51)
  * ``run()``    = 25 angles 0..360 step 15, total 0 -> 1, p = count/total   (Matcher.py:305-338)
  * ``optRun()`` = +-30 degree window, unmatched images count 1             (Matcher.py:340-388)
  * the positional-argument quirk Matcher('SIFT', 320, 240) -> index=320, width=240 (Matcher.py:41).

What differs, on purpose
  * the query is extracted once per ``setQuery`` instead of once per map image (the reference re-runs
    detectAndCompute inside every *Match call, Matcher.py:177,205,242,270; extraction is deterministic so
    results are unchanged);
  * SIFT/SURF use exact search (what cv2.BFMatcher.knnMatch(k=2) returns) instead of FLANN's approximate
    KD-forest (Matcher.py:212-218,273-279);
  * ORBMatch in the reference raises NameError (Matcher.py:194).  ``orb_mode='knn'`` (default) is 2-NN +
    ratio test; ``orb_mode='crosscheck'`` is the mutual-nearest-neighbour count the authors' older bytecode ran;
  * an image (or query) with fewer than 2 descriptors counts 0 instead of raising;
  * 'Color': histograms come from cv2.calcHist / cv2.normalize exactly as in the reference (Matcher.py:67-85); the
    chi-squared search of a query against a whole colour index (search.py:7-26) is one GPU call (mcl_chi2) whose
    float32 results equal the reference's bit for bit, so ``colorSearch`` / ``run`` return the same numbers.
'''

import glob

import cv2
import numpy as np

from . import _lib

extension = '.png'

_EXTRACTOR_ALG = {'SIFT': 'SIFT', 'SURF': 'SURF', 'ORB': 'ORB', 'BOW': 'SIFT'}


def _make_extractor(alg):
    if alg == 'SIFT':
        x = getattr(cv2, 'xfeatures2d', None)
        return x.SIFT_create() if x is not None and hasattr(x, 'SIFT_create') else cv2.SIFT_create()
    if alg == 'SURF':
        x = getattr(cv2, 'xfeatures2d', None)
        if x is None or not hasattr(x, 'SURF_create'):
            raise RuntimeError("cv2.xfeatures2d.SURF_create is unavailable in this OpenCV build (non-free); "
                               "pass SURF descriptors with setQueryDescriptors()/setIndex()")
        return x.SURF_create()
    return cv2.ORB_create()


class _ResidentCache(object):
    """id(index dict) -> resident MapIndex.  Holds a reference to the dict so the id stays valid, and a fingerprint of its
    content -- the paths in order and the identity + shape of every descriptor array -- so that an index changed in place
    (index[path] = (kp, other_des), a path added or removed) gets a fresh HBM copy instead of a stale one."""

    def __init__(self, capacity=64):
        self.capacity = capacity
        self.entries = {}
        self.order = []

    @staticmethod
    def _fingerprint(index):
        return tuple((p, id(v[1]), None if v[1] is None else v[1].shape) for p, v in index.items())

    def get(self, ctx, kind, index):
        key = (id(index), kind)
        fp = self._fingerprint(index)
        ent = self.entries.get(key)
        if ent is not None and ent[0] is index and ent[3] == fp:
            return ent[1], ent[2]
        if ent is not None:          # same dict object, different content: release the stale resident copy
            ent[1].close()
            self.order.remove(key)
        paths = list(index.keys())
        descs = [index[p][1] for p in paths]
        resident = _lib.MapIndex(ctx, kind, descs)
        col = dict((p, i) for i, p in enumerate(paths))
        self.entries[key] = (index, resident, col, fp)
        self.order.append(key)
        while len(self.order) > self.capacity:
            old = self.order.pop(0)
            e = self.entries.pop(old, None)
            if e is not None:
                e[1].close()
        return resident, col


class _BowCache(object):
    """(index path, mtime, size) -> resident BowIndex, least recently used first out: a 10 000-word index keeps ~9 MB in
    HBM (code book + planes + tf-idf rows), and the reference's loop visits every location's .pkl for every frame."""

    def __init__(self, capacity=512):
        from collections import OrderedDict
        self.capacity = capacity
        self.entries = OrderedDict()

    def get(self, key):
        ent = self.entries.get(key)
        if ent is not None:
            self.entries.move_to_end(key)
        return ent

    def put(self, key, ent):
        self.entries[key] = ent
        while len(self.entries) > self.capacity:
            _, old = self.entries.popitem(last=False)
            old.close()


_resident_cache = _ResidentCache()
_bow_cache = _BowCache()


class Matcher(object):

    ######################
    ### Initialization ###
    ######################

    def __init__(self, algorithm, index=None, width=800, height=600, *, imagesPerLocation=25, angleStep=15):
        """Reference signature (Matcher.py:41) plus two keyword-only generalisations of its hard-coded map shape
        (`range(0, 375, 15)`, Matcher.py:309,325,348): images per location and degrees between them."""
        self.w = width
        self.h = height
        self.alg = algorithm
        self.index = index
        self.imagesPerLocation = int(imagesPerLocation)
        self.angleStep = int(angleStep)
        self.numWords = 10000
        self.orb_mode = 'knn'
        self.ratio = 0.7
        self._qdes = {}
        self._counts = None
        self._ctx = None

    def _context(self):
        if self._ctx is None:
            self._ctx = _lib.default_context()
        return self._ctx

    def setQuery(self, imagePath):
        self.image = cv2.imread(imagePath)
        self.image = cv2.bilateralFilter(self.image, 9, 75, 75)
        self.image = cv2.resize(self.image, (self.w, self.h))
        self._qdes = {}
        self._counts = None

    def setQueryDescriptors(self, descriptors, algorithm=None):
        """Descriptor-level entry (no reference counterpart): use these rows as the query's extractor output."""
        self._qdes = {_EXTRACTOR_ALG[algorithm or self.alg]: descriptors}
        self._counts = None

    def setDirectory(self, directory):
        self.data = directory

    def setIndex(self, index):
        self.index = index
        self._counts = None

    def setColorIndex(self, colorIndex):
        self.colorIndex = colorIndex

    ############################
    ### Color-Based Matching ###
    ############################

    def createHistogram(self, image, bins=[8, 8, 8]):
        '''
        Creates a flattened 3D histogram (reference Matcher.py:67-74; stays on cv2 so the inputs are identical).
        '''
        hist = cv2.calcHist([image], [0, 1, 2], None, bins, [0, 256, 0, 256, 0, 256])
        hist = cv2.normalize(hist, hist, 0, 255, cv2.NORM_MINMAX)
        return hist.flatten()

    def createColorIndex(self):
        '''
        Creates dictionary with keys as image names and histograms as values (reference Matcher.py:76-88).
        '''
        index = {}
        for imagePath in glob.glob(self.data + "/*" + extension):
            filename = imagePath[imagePath.rfind("/") + 1:]
            image = cv2.imread(imagePath)
            index[filename] = self.createHistogram(image)
        return index

    def colorSearch(self):
        '''
        Searches query image against index; results are (chi-squared distance, image name) sorted ascending
        (reference Matcher.py:89-99 + search.py:7-26).  All distances of the index come from one mcl_chi2 call.
        '''
        names = list(self.colorIndex.keys())
        queryFeatures = self.createHistogram(self.image)
        if not names:
            return []
        d = self._context().chi2(queryFeatures[None, :], np.stack([self.colorIndex[k] for k in names]))[0]
        return sorted([(v, k) for (k, v) in zip(names, d)])

    def _color_finish(self, results):
        """The colour branch of run()/optRun() (reference Matcher.py:328-336, :377-385): distances are inverted into
        weights and normalised.  The reference does this with np.float32 scalars, summing in ascending-distance order;
        the same operations in the same order are kept so that the float32 roundings agree."""
        chi = [d for d, _ in results]
        totalMatches = 300000. / sum(chi)
        inverse = [200. / d for d in chi]
        norm = sum(inverse)
        weight = {}
        for (_, name), w in zip(results, inverse):
            weight[name] = w / norm * totalMatches
        by_angle = sorted(weight, key=lambda name: int(name.replace(extension, '').replace('angle', '')))
        return totalMatches, [weight[name] / totalMatches for name in by_angle]

    #############################
    ### Bag-of-Words Matching ###
    #############################

    def createIndex(self, trainingPath):
        """Reference Matcher.py:106-141.  kmeans(train_descriptors, numWords, 1) = one run of Lloyd iterations from numWords
        observations drawn with numpy's global RandomState (scipy's _kpoints, unseeded unless the caller seeds
        np.random); the iterations (assignment on tensor cores, means, empty clusters dropped, distortion threshold 1e-5)
        run on the GPU (mcl_kmeans), as does the visual-word assignment of the training descriptors."""
        import joblib
        from sklearn import preprocessing
        desc = _make_extractor('SIFT')
        train_des_list = []
        numWords = self.numWords
        image_paths = glob.glob(trainingPath + '/*' + '.png')
        for imagePath in image_paths:
            print(imagePath)
            image = cv2.imread(imagePath)
            kp, des = desc.detectAndCompute(image, None)
            train_des_list.append((imagePath, des))
        train_descriptors = np.vstack([d for _, d in train_des_list if d is not None])
        ctx = self._context()
        idx = np.random.mtrand._rand.choice(train_descriptors.shape[0], size=int(numWords), replace=False)
        voc, variance, _ = ctx.kmeans(train_descriptors, np.take(train_descriptors, idx, axis=0))
        voc = np.ascontiguousarray(voc, dtype=np.float32)
        tmp = _lib.BowIndex(ctx, [(np.zeros((0, numWords), 'float32'), np.ones(numWords, 'float32'), voc)], numWords)
        _, words = tmp.score([d for _, d in train_des_list], want_words=True)
        tmp.close()
        im_features = np.zeros((len(image_paths), numWords), "float32")
        pos = 0
        for i, (_, d) in enumerate(train_des_list):
            n = 0 if d is None else len(d)
            np.add.at(im_features[i], words[0, pos:pos + n], 1)
            pos += n
        nbr_occurences = np.sum((im_features > 0) * 1, axis=0)
        idf = np.array(np.log((1.0 * len(image_paths) + 1) / (1.0 * nbr_occurences + 1)), 'float32')
        im_features = im_features * idf
        im_features = preprocessing.normalize(im_features, norm='l2')
        joblib.dump((im_features, image_paths, idf, numWords, voc), trainingPath + ".pkl", compress=3)

    def writeIndices(self):
        """Reference Matcher.py:143-145.  One process per GPU: the locations are dealt round-robin to the ranks (each index is
        an independent k-means + histogram pass writing its own .pkl), then the ranks meet at a barrier."""
        from .engine import _dist
        d = _dist()
        world, rank = (d.get_world_size(), d.get_rank()) if d is not None else (1, 0)
        for mapp in sorted(glob.glob('map/*/'))[rank::world] if world > 1 else glob.glob('map/*/'):
            self.createIndex(mapp[:-1])
        if world > 1:
            d.barrier()

    def createFeatureIndex(self):
        '''
        Creates a dictionary with keys as image paths and values as keypoints and descriptors
        (reference Matcher.py:147-163; extraction stays on cv2 so the descriptors are identical).
        '''
        import os
        from concurrent.futures import ThreadPoolExecutor
        alg = self.alg if self.alg in ('SURF', 'SIFT') else 'ORB'
        _make_extractor(alg)  # raise here (not in a worker) if the extractor is unavailable
        paths = glob.glob(self.data + '/*' + extension)

        def one(imagePath):
            # a fresh extractor per image: cv2 extractors are deterministic across instances and threads,
            # and cv2 releases the GIL, so the map images are extracted in parallel; dict order = glob order
            image = cv2.imread(imagePath)
            return _make_extractor(alg).detectAndCompute(image, None)

        workers = max(1, min(len(paths), int(os.environ.get("MCL_EXTRACT_THREADS", os.cpu_count() or 1))))
        if workers == 1:
            results = [one(p) for p in paths]
        else:
            with ThreadPoolExecutor(max_workers=workers) as pool:
                results = list(pool.map(one, paths))
        index = {}
        for imagePath, (kp, des) in zip(paths, results):
            index[imagePath] = (kp, des)
        return index

    #################################
    ### Image Matching Algorithms ###
    #################################

    def _query_descriptors(self, alg):
        if alg not in self._qdes:
            kp, des = _make_extractor(alg).detectAndCompute(self.image, None)
            self._qdes[alg] = des
        return self._qdes[alg]

    def _mode(self, kind):
        if kind == 'ORB' and self.orb_mode == 'crosscheck':
            return _lib.CROSSCHECK
        return _lib.KNN2_RATIO

    def _match_index(self, kind, only_paths=None):
        """counts of the current query against every image of self.index (one launch), cached per query.
        only_paths: restrict the work to these images (DOR); the others are not matched."""
        if not self.index:
            raise ValueError("no feature index set (setIndex / createFeatureIndex)")
        ctx = self._context()
        resident, col = _resident_cache.get(ctx, kind, self.index)
        des = self._query_descriptors(kind)
        if only_paths is None:
            key = (id(self.index), kind, self._mode(kind), float(self.ratio), id(resident))
            if self._counts is not None and self._counts[0] == key:
                return self._counts[1], col
            counts = resident.match_counts([des], mode=self._mode(kind), ratio=self.ratio)[0]
            self._counts = (key, counts)
            return counts, col
        mask = np.zeros((1, resident.n_images), dtype=np.uint8)
        for p in only_paths:
            mask[0, col[p]] = 1
        counts = resident.match_counts([des], mode=self._mode(kind), ratio=self.ratio, mask=mask, fill_value=1)[0]
        return counts, col

    def _single(self, kind, imagePath):
        counts, col = self._match_index(kind)
        return int(counts[col[imagePath]])

    def ORBMatch(self, imagePath, display_results=False):
        '''
        Hamming brute force between the query's ORB descriptors and one indexed image (reference
        Matcher.py:169-194).  orb_mode 'knn': 2-NN + ratio test; 'crosscheck': mutual nearest neighbours.
        '''
        return self._single('ORB', imagePath)

    def SURFMatch(self, imagePath, display_results=False):
        '''
        Exact 2-NN under L2 on 64-d float descriptors + Lowe's ratio test (reference Matcher.py:196-230).
        '''
        if not self.index:
            surf = _make_extractor('SURF')
            kp2, des2 = surf.detectAndCompute(cv2.imread(imagePath), None)
            tmp = _lib.MapIndex(self._context(), 'SURF', [des2])
            n = int(tmp.match_counts([self._query_descriptors('SURF')], ratio=self.ratio)[0, 0])
            tmp.close()
            return n
        return self._single('SURF', imagePath)

    def BOWMatch(self, indexPath):
        '''the query's score against an individual index (reference Matcher.py:232-259)'''
        import joblib
        import os
        st = os.stat(indexPath)
        key = (os.path.abspath(indexPath), st.st_mtime_ns, st.st_size)
        ent = _bow_cache.get(key)
        if ent is None:
            im_features, image_paths, idf, numWords, voc = joblib.load(indexPath)
            # the reference sizes the query histogram with self.numWords (Matcher.py:237); it must equal the index width
            W = int(np.asarray(im_features).shape[1])
            ent = _lib.BowIndex(self._context(), [(im_features, idf, voc)], W)
            _bow_cache.put(key, ent)
        des = self._query_descriptors('SIFT')
        scores, _ = ent.score([des])
        return scores

    def SIFTMatch(self, imagePath, display_results=False):
        '''
        Exact 2-NN under L2 on 128-d SIFT descriptors + Lowe's ratio test (reference Matcher.py:262-291).
        '''
        return self._single('SIFT', imagePath)

    def _angle_paths(self):
        return [self.data + '/angle' + str(i).zfill(3) + extension
                for i in range(0, self.imagesPerLocation * self.angleStep, self.angleStep)]

    def _finish(self, numbers):
        totalMatches = sum(numbers)
        if totalMatches == 0:
            totalMatches = 1
        return totalMatches, list(map(lambda x: x / totalMatches, numbers))

    def run(self):
        if self.alg != 'Color' and self.alg != 'BOW':
            kind = self.alg if self.alg in ('SIFT', 'SURF') else 'ORB'
            counts, col = self._match_index(kind)
            return self._finish([int(counts[col[p]]) for p in self._angle_paths()])
        elif self.alg == 'BOW':
            score = self.BOWMatch(self.data + '.pkl')
            return 10 * np.max(score), score[0].tolist()
        return self._color_finish(self.colorSearch())

    def optRun(self, bestAngleIndex):
        if bestAngleIndex is None:
            return self.run()
        if self.alg == 'Color' or self.alg == 'BOW':
            # the reference takes the colour branch for both here (Matcher.py:377-385)
            return self._color_finish(self.colorSearch())
        kind = self.alg if self.alg in ('SIFT', 'SURF') else 'ORB'
        window = angle_window(bestAngleIndex, self.imagesPerLocation, self.angleStep)
        paths = self._angle_paths()
        counts, col = self._match_index(kind, only_paths=[p for p, w in zip(paths, window) if w])
        return self._finish([int(counts[col[p]]) if w else 1 for p, w in zip(paths, window)])


def angle_window(bestAngleIndex, n_images=25, step=15):
    """Which of the angle images optRun matches (reference Matcher.py:342-370): +-30 degrees around the best
    angle, with the reference's wrap-around branch (modulo 360 on both bounds)."""
    bestAngle = bestAngleIndex * step
    lower = bestAngle - 30
    upper = bestAngle + 30
    full = (n_images - 1) * step
    out = []
    for k in range(n_images):
        i = k * step
        if lower >= 0 and upper <= full:
            out.append(i >= lower and i <= upper)
        else:
            out.append(i >= lower % full or i <= upper % full)
    return out


if __name__ == '__main__':

    print(__doc__)
