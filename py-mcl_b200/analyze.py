'''
Visual Monte Carlo Localization on B200
=======================================

Drop-in for the reference's ``analyzer`` (reference analyze.py:32-445): same constructor, same
``createIndex/createRawP/processRaw/optP`` drivers, same ``rawP.txt`` / ``out.txt`` / ``bestGuess.txt`` text
formats, same filter arithmetic (float64, the reference's operation order, including its quirks: the
degrees->radians slip in the forward factor, ``1/75`` priors for 25 angles, and the aliasing by which
``accountCommand`` of frame t+1 rewrites the lists already stored for frame t).

What runs on the GPU (libmcl.so): all descriptor matching -- for ``createRawP`` every frame x every location x
every angle in ONE batched call (map images sharded over the GPUs of the box, int32 counts all-gathered) --
the bag-of-words assignment + tf-idf scoring, the variance-of-Laplacian blur weight, and the particle
shift / reweight update.  Keypoint extraction and image decode stay on cv2 so that inputs are identical.
'''

import glob
import os
import time

import cv2
import numpy as np

from . import _lib
from .engine import ShardedBow, ShardedMap
from .Matcher import Matcher, angle_window

extension = '.png'


class analyzer(object):

    def __init__(self, method, width, height, numLocations=7, *, imagesPerLocation=25, angleStep=15, framesPerCall=256):
        """Reference signature (analyze.py:34) plus keyword-only generalisations of its hard-coded map shape: numLocations
        (7, analyze.py:35), images per location (25) and degrees between them (15) (`range(0, 375, 15)`, Matcher.py:309;
        `[1/75]*25`, analyze.py:107).  framesPerCall bounds how many frames createRawP extracts and matches at a time."""
        self.numLocations = numLocations
        self.I = int(imagesPerLocation)
        self.angleStep = int(angleStep)
        self.framesPerCall = max(1, int(framesPerCall))
        self.indices = [None] * self.numLocations
        self.method = method
        self.w = width
        self.h = height
        self.rawP = []
        self.blurP = []
        self.commands = self.readCommand('commands.txt')
        self.bestGuess = []
        self.orb_mode = 'knn'
        # optional binary side-car of rawP.txt (rawP.npz): processRaw then skips the text parse
        self.sidecar = bool(os.environ.get('MCL_SIDECAR'))
        self._map = None
        self._ctx = None

    def _context(self):
        if self._ctx is None:
            self._ctx = _lib.default_context()
        return self._ctx

    def createIndex(self):
        """
        Create the feature indices (reference analyze.py:45-58).
        """
        matcher = self._matcher()
        if self.method != 'BOW':
            for i in range(self.numLocations):
                matcher.setDirectory('map/' + str(i))
                if self.method != 'Color':
                    self.indices[i] = matcher.createFeatureIndex()
                else:
                    self.indices[i] = matcher.createColorIndex()
        else:
            matcher.writeIndices()
        self._map = None

    ####################
    ### Main Methods ###
    ####################

    def _matcher(self):
        return Matcher(self.method, width=self.w, height=self.h, imagesPerLocation=self.I, angleStep=self.angleStep)

    def _prior(self):
        # the reference's [1, [1/75]*25] (analyze.py:107,166,204): 1/(3 I) per image
        return [1, [1 / (3 * self.I)] * self.I]

    def _kind(self):
        return self.method if self.method in ('SIFT', 'SURF') else 'ORB'

    def _mode(self):
        return _lib.CROSSCHECK if (self._kind() == 'ORB' and self.orb_mode == 'crosscheck') else _lib.KNN2_RATIO

    def _angle_paths(self, loc):
        return ['map/' + str(loc) + '/angle' + str(i).zfill(3) + extension for i in range(0, self.I * self.angleStep, self.angleStep)]

    def _resident_map(self):
        """All locations x 25 angle images as one sharded, HBM-resident map (run()'s image order)."""
        if self._map is None:
            descs = []
            for loc in range(self.numLocations):
                for p in self._angle_paths(loc):
                    descs.append(self.indices[loc][p][1])
            self._map = ShardedMap(self._kind(), descs, ctx=self._context())
        return self._map

    def _extract(self, matcher, imagePath):
        matcher.setQuery(imagePath)
        return matcher._query_descriptors(self._kind())

    def _extract_many(self, imagePaths, kind=None):
        """Query descriptors of many frames.  setQuery + detectAndCompute per frame exactly as the reference does them
        (Matcher.py:48-51 + the extractor call of *Match), on a thread pool: cv2 releases the GIL, extraction is
        deterministic and independent per frame, and it is the end-to-end bottleneck once matching runs on the GPU
        (SURVEY D5 / hard part 7).  Order is preserved."""
        kind = kind or self._kind()

        def one(path):
            m = self._matcher()
            m.setQuery(path)
            return m._query_descriptors(kind)

        return self._per_frame(one, imagePaths)

    def _per_frame(self, one, imagePaths):
        """[one(path) for path in imagePaths] on a thread pool, the frames dealt round-robin to the ranks of the job."""
        import os
        from concurrent.futures import ThreadPoolExecutor

        def many(paths):
            workers = max(1, min(len(paths), int(os.environ.get("MCL_EXTRACT_THREADS", os.cpu_count() or 1))))
            if workers <= 1:
                return [one(p) for p in paths]
            with ThreadPoolExecutor(max_workers=workers) as pool:
                return list(pool.map(one, paths))

        # one process per GPU: the frames are dealt to the ranks, each extracts its share, the results are all-gathered
        from .engine import _dist
        d = _dist()
        if d is None or d.get_world_size() == 1:
            return many(imagePaths)
        world, rank = d.get_world_size(), d.get_rank()
        parts = [None] * world
        d.all_gather_object(parts, many(imagePaths[rank::world]))
        out = [None] * len(imagePaths)
        for r in range(world):
            out[r::world] = parts[r]
        return out

    @staticmethod
    def _run_result(counts):
        numbers = [int(c) for c in counts]
        totalMatches = sum(numbers)
        if totalMatches == 0:
            totalMatches = 1
        return [totalMatches, list(map(lambda x: x / totalMatches, numbers))]

    def createRawP(self):
        """
        Raw probabilities directly from image matching, stored in rawP.txt (reference analyze.py:64-96).
        All frames x all map images are matched in one batched GPU call.
        """
        if self.method != 'BOW':
            print('Creating indices...')
            self.createIndex()

        start = time.time()
        p = []
        matcher = self._matcher()
        print('Matching...')
        frames = glob.glob('cam1_img' + '/*' + extension)
        I = self.I

        if self.method == 'Color':
            # every frame's histogram against every location's colour index: one chi-squared call
            def hist(path):
                m = self._matcher()
                m.setQuery(path)
                return m.createHistogram(m.image)

            qh = self._per_frame(hist, frames)
            every = list(range(self.numLocations))
            per_frame = self._color_locations(matcher, qh, every) if frames else []
            for imagePath, results in zip(frames, per_frame):
                p.extend(results)
                print('\t' + imagePath)
        elif self.method != 'BOW':
            # frames in bounded chunks (the reference streams frame by frame; one call per chunk keeps the batching win
            # while host memory, the pinned staging copy and the device buffers stop growing with the run length)
            resident = self._resident_map()
            chunks = []
            for c0 in range(0, len(frames), self.framesPerCall):
                qdes = self._extract_many(frames[c0:c0 + self.framesPerCall])
                chunks.append(resident.match_counts(qdes, mode=self._mode()))
            counts = np.concatenate(chunks, axis=0) if chunks else np.zeros((0, self.numLocations * I), np.int32)
            for f, imagePath in enumerate(frames):
                for i in range(self.numLocations):
                    p.append(self._run_result(counts[f, i * I:(i + 1) * I]))
                print('\t' + imagePath)
        else:
            import joblib

            def load(i):
                im_features, image_paths, idf, numWords, voc = joblib.load('map/' + str(i) + '.pkl')
                return im_features, idf, voc

            # locations (code book + index) are the unit of sharding; the float32 scores are all-gathered
            bow = ShardedBow(self.numLocations, load, ctx=self._context())
            parts = []
            for c0 in range(0, len(frames), self.framesPerCall):
                parts.append(bow.score(self._extract_many(frames[c0:c0 + self.framesPerCall], 'SIFT')))
            scores = np.concatenate(parts, axis=0) if parts else np.zeros((0, bow.total_images), np.float32)
            for f, imagePath in enumerate(frames):
                for i in range(self.numLocations):
                    s = scores[f:f + 1, bow.img_off[i]:bow.img_off[i + 1]]
                    p.append([10 * np.max(s), s[0].tolist()])
                print('\t' + imagePath)
            bow.close()
        self.rawP = p
        self.writeProb(p, 'rawP.txt', 'w')
        if self.sidecar:
            self.writeSidecar(p, 'rawP.txt', counts if self.method not in ('BOW', 'Color') else None)

        end = time.time()
        print('Time elapsed: %0.1f' % (end - start))

    def _color_locations(self, matcher, query_hists, locs):
        """For each query histogram, run()'s [total, probs] of every location in `locs` (the others get the DOR prior):
        the chi-squared distances to all those locations' colour indices come from one mcl_chi2 call."""
        names = dict((i, list(self.indices[i].keys())) for i in locs)
        stacked = [self.indices[i][k] for i in locs for k in names[i]]
        chi = self._context().chi2(np.stack(query_hists), np.stack(stacked))
        out = []
        for f in range(len(query_hists)):
            results, col = [], 0
            for i in range(self.numLocations):
                if i not in names:
                    results.append(self._prior())
                    continue
                matcher.setDirectory('map/' + str(i))
                d = chi[f, col:col + len(names[i])]
                col += len(names[i])
                totalMatches, probL = matcher._color_finish(sorted([(v, k) for (k, v) in zip(names[i], d)]))
                results.append([totalMatches, probL])
            out.append(results)
        return out

    def _filter(self, command, previousProbs, p, blurFactor):
        """accountCommand -> prevWeight -> probUpdate -> best guess in one kernel (mcl_filter_step).
        Reproduces the reference's aliasing: the command shift is written back into the SAME inner lists of
        previousProbs (analyze.py:317-335 mutate a shallow copy)."""
        prev2, adjusted, best = self._stage(_lib.F_ALL, previousProbs, p, blurFactor, command)
        self._write_back(command, previousProbs, prev2)
        return adjusted, best

    @staticmethod
    def _write_back(command, lists, shifted):
        """accountCommand's in-place effects on the caller's lists: rotated probability lists, one bumped total."""
        if command == 'l' or command == 'r':
            for circles, row in zip(lists, shifted):
                circles[1] = row[1]
        elif command == 'f':
            for circles, row in zip(lists, shifted):
                if row[0] != circles[0]:
                    circles[0] = row[0]

    def processRaw(self):
        """
        Processes the raw data from rawP.txt (reference analyze.py:98-146).
        """
        previousProbs = []
        for i in range(self.numLocations):
            previousProbs.append(self._prior())

        start = time.time()
        probDict = self.readSidecar('rawP.txt') if self.sidecar else None
        if probDict is None:
            probDict = self.readProb('rawP.txt')
        blurP = []
        frames = glob.glob('cam1_img' + '/*' + extension)
        blurs = self._laplacians(frames)

        for imagePath, blurFactor in zip(frames, blurs):
            stem = imagePath.replace('cam1_img/', '').replace(extension, '')
            p = probDict[stem]
            command = self.commands[stem]
            adjusted, best = self._filter(command, previousProbs, p, blurFactor)
            self.bestGuess.extend([[best[0], best[1]]])
            blurP.extend(adjusted)
            previousProbs = adjusted
            print(imagePath)

        self.blurP = blurP
        self.writeProb(self.blurP, 'out.txt', 'w')
        self.writeProb(self.bestGuess, 'bestGuess.txt', 'w')

        end = time.time()
        print('Time elapsed: %0.1f' % (end - start))

    def optP(self):
        """
        Dynamically Optimized Retrieval (reference analyze.py:148-232): after the first frame only locations
        within +-2 of the best location are matched, and inside them only angles within +-30 degrees; every
        other particle gets a small non-zero weight.  Frame t+1's window depends on frame t's best guess, so
        this is one masked GPU call per frame (the frames' descriptors and blur weights are prepared up front).
        """
        if self.method != 'BOW':
            print('Creating indices...')
            self.createIndex()

        blurP = []
        previousProbs = []
        bestAngleIndex = None
        bestCircleIndex = None

        for i in range(self.numLocations):
            previousProbs.append(self._prior())

        matcher = self._matcher()
        start = time.time()
        print('Matching...')
        color = self.method == 'Color'
        resident = None if color else self._resident_map()
        I = self.I
        n_img = self.numLocations * I

        # Which images a frame is matched against depends on the previous frame's best guess, but its descriptors
        # (histogram) and its blur weight do not: extract / decode every frame up front on the thread pool (cv2 extraction
        # is the sequential cost of this driver otherwise), then run the dependent chain of masked GPU calls.
        frames = glob.glob('cam1_img' + '/*' + extension)
        if color:
            def hist(path):
                m = self._matcher()
                m.setQuery(path)
                return m.createHistogram(m.image)

            features = self._per_frame(hist, frames)
        else:
            features = self._extract_many(frames)
        blurs = self._laplacians(frames)

        for imagePath, feature, blurFactor in zip(frames, features, blurs):
            if bestCircleIndex is None:
                locs = list(range(self.numLocations))
            else:
                lower = bestCircleIndex - 2
                upper = bestCircleIndex + 2
                locs = [i for i in range(self.numLocations) if i >= lower and i <= upper]
            if color:
                # the colour branch of optRun ignores the angle window (reference Matcher.py:377-385)
                p = self._color_locations(matcher, [feature], locs)[0]
            else:
                des = feature
                window = angle_window(bestAngleIndex, I, self.angleStep) if bestAngleIndex is not None else [True] * I
                mask = np.zeros((1, n_img), dtype=np.uint8)
                for i in locs:
                    mask[0, i * I:(i + 1) * I] = window
                counts = resident.match_counts([des], mode=self._mode(), mask=mask, fill_value=1)[0]
                p = []
                for i in range(self.numLocations):
                    if i in locs:
                        p.append(self._run_result(counts[i * I:(i + 1) * I]))
                    else:
                        p.append(self._prior())
            print('\t' + imagePath)

            command = self.commands[imagePath.replace('cam1_img/', '').replace(extension, '')]
            adjusted, best = self._filter(command, previousProbs, p, blurFactor)
            bestCircleIndex, bestAngleIndex = best
            self.bestGuess.extend([[bestCircleIndex, bestAngleIndex]])
            blurP.extend(adjusted)
            previousProbs = adjusted

        self.blurP = blurP
        self.writeProb(self.blurP, 'out.txt', 'w')
        self.writeProb(self.bestGuess, 'bestGuess.txt', 'w')
        end = time.time()
        print('Time elapsed: %0.1f' % (end - start))

    ############################
    ### Probability Updating ###
    ############################

    @staticmethod
    def _tag(x):
        if isinstance(x, np.float32):
            return _lib.T_F32
        if isinstance(x, np.float64):
            return _lib.T_F64
        return _lib.T_PY

    @classmethod
    def _pack(cls, lists):
        """[[total, [p...]]...] -> (L, 1+I) float64 values + (L, 2) scalar-type tags [total, probabilities]."""
        values = np.array([[c[0]] + list(c[1]) for c in lists], dtype=np.float64)
        tags = np.array([[cls._tag(c[0]), cls._tag(c[1][0]) if len(c[1]) else _lib.T_PY] for c in lists], dtype=np.int8)
        return values, tags

    @staticmethod
    def _unpack(values, tags):
        """The inverse of _pack: Python floats, np.float64 or np.float32 scalars according to the tags."""
        out = []
        for row, (t0, t1) in zip(values, tags):
            total = float(row[0]) if t0 == _lib.T_PY else (np.float32(row[0]) if t0 == _lib.T_F32 else np.float64(row[0]))
            if t1 == _lib.T_PY:
                probs = row[1:].tolist()
            elif t1 == _lib.T_F32:
                probs = list(row[1:].astype(np.float32))
            else:
                probs = list(row[1:])
            out.append([total, probs])
        return out

    def _stage(self, stages, previousP, currentP, blurFactor=0.0, command='s'):
        """Runs the selected filter stages on the GPU.  Lists that hold only Python numbers take mcl_filter_step
        (float64); lists with np.float32 values (colour method) take mcl_filter_step_typed, which applies numpy's scalar
        promotion per operation like the reference's Python operators do.
        Returns (previousP after the command, result lists, (bestCircle, bestAngle))."""
        prev, pt = self._pack(previousP)
        cur, ct = self._pack(currentP)
        if not (pt == _lib.T_F32).any() and not (ct == _lib.T_F32).any():
            prev2, out, best = self._context().filter_step(stages, command, float(blurFactor), prev, cur, angle_step=self.angleStep)
            return self._unpack(prev2, pt), [[float(r[0]), r[1:].tolist()] for r in out], best
        prev2, out, ot, best = self._context().filter_step_typed(stages, command, float(blurFactor), self._tag(blurFactor),
                                                                  prev, pt, cur, ct, angle_step=self.angleStep)
        return self._unpack(prev2, pt), self._unpack(out, ot), best

    def probUpdate(self, previousP, currentP, blurFactor):
        """Blur-proportional blend (reference analyze.py:239-276)."""
        return self._stage(_lib.F_BLUR, previousP, currentP, blurFactor)[1]

    def prevWeight(self, previousP, currentP):
        """0.7 / 0.3 blend (reference analyze.py:278-310)."""
        return self._stage(_lib.F_PREV, previousP, currentP)[1]

    def accountCommand(self, command, previousP):
        """Shift particles according to the current command (reference analyze.py:312-336); mutates the inner
        lists of previousP like the reference's shallow copy does."""
        prev2 = self._stage(_lib.F_COMMAND, previousP, previousP, 0.0, command)[0]
        copy = previousP[:]
        self._write_back(command, copy, prev2)
        return copy

    ###################################
    ### Reading and Writing Methods ###
    ###################################

    def writeProb(self, prob, filename, mode):
        ''' write out the probabilistic values to a txt file (reference analyze.py:351-356) '''
        from .engine import _dist
        d = _dist()
        if d is not None and d.get_rank() != 0:
            return   # one process per GPU: every rank holds the same lists, rank 0 writes the files
        file = open(filename, mode)
        # np.float32 list elements (colour method) must print as plain numbers or readProb cannot parse them back
        with np.printoptions(legacy='1.25'):
            for index in prob:
                file.write(str(index[0]) + '\n')
                file.write(str(index[1]) + '\n')
        file.close()

    @staticmethod
    def _text_digest(filename):
        import hashlib
        with open(filename, 'rb') as f:
            return hashlib.sha256(f.read()).hexdigest()

    def writeSidecar(self, prob, filename, counts=None):
        """Binary twin of a probability file (no reference counterpart): <stem>.npz with the float64 values readProb
        would parse out of the text (float(str(x)) of every number written), the raw int32 match counts when the method
        has them, and the SHA-256 of the text file it mirrors.  Written by rank 0 only, like the text."""
        from .engine import _dist
        d = _dist()
        if d is not None and d.get_rank() != 0:
            return
        L = self.numLocations
        with np.printoptions(legacy='1.25'):
            total = np.array([float(str(c[0])) for c in prob], dtype=np.float64).reshape(-1, L)
            flat = np.array([float(str(x)) for c in prob for x in c[1]], dtype=np.float64)
        # per (frame, location) list lengths: equal for the feature methods, ragged for BOW (locations own different image counts)
        lengths = np.array([len(c[1]) for c in prob], dtype=np.int64).reshape(-1, L)
        extra = {} if counts is None else {'counts': np.asarray(counts, dtype=np.int32)}
        np.savez(filename[:filename.rfind('.')] + '.npz', total=total, probs=flat, lengths=lengths,
                 text_sha256=np.array(self._text_digest(filename)), **extra)

    def readSidecar(self, filename):
        """{frame stem: [[total, [p...]] * numLocations]} from <stem>.npz -- the same dictionary readProb(filename) builds --
        or None when there is no side-car or it does not belong to the text file on disk."""
        path = filename[:filename.rfind('.')] + '.npz'
        if not (os.path.isfile(path) and os.path.isfile(filename)):
            return None
        with np.load(path) as z:
            if str(z['text_sha256']) != self._text_digest(filename) or z['total'].shape[1] != self.numLocations:
                return None
            total, flat, lengths = z['total'], z['probs'], z['lengths']
        probD = {}
        pos = 0
        for f in range(total.shape[0]):
            row = []
            for l in range(total.shape[1]):
                n = int(lengths[f, l])
                row.append([float(total[f, l]), flat[pos:pos + n].tolist()])
                pos += n
            probD[str(f).zfill(4)] = row
        return probD

    def readCommand(self, filename):
        '''reads the command list from the robot (reference analyze.py:358-365)'''
        file = open(filename, 'r')
        content = file.read().split('\n')[:-1]
        file.close()
        commandDict = {}
        for data in content:
            commandDict[data[:4]] = str(data[-1])
        return commandDict

    def readBestGuess(self, filename):
        '''reads the list of best guesses (reference analyze.py:367-373)'''
        file = open(filename, 'r')
        content = file.read().split('\n')[:-1]
        file.close()
        content = list(map(int, content))
        bestGuesses = [[content[x], content[x+1]] for x in range(len(content) - 1)[::2]]
        return bestGuesses

    def readCoord(self, filename):
        file = open(filename, 'r')
        content = file.read().split('\n')[:-1]
        file.close()
        coordinates = [list(map(int, coord.split(','))) for coord in content]
        return coordinates

    def readProb(self, filename):
        '''reads a probability file into {frame stem: [[total, [p...]] * numLocations]} (reference analyze.py:381-398)'''
        file = open(filename, 'r')
        raw_content = file.read().split('\n')[:-1]
        file.close()
        raw_chunks = [raw_content[i:i+2] for i in range(0, len(raw_content), 2)]
        raw_probL = [raw_chunks[i:i+self.numLocations] for i in range(0, len(raw_chunks), self.numLocations)]
        probD = {}
        counter = 0
        for prob in raw_probL:
            content = []
            for location in prob:
                totalMatches = float(location[0])
                probabilities = list(map(float, location[1].replace('[', '').replace(']', '').split(',')))
                content.append([totalMatches, probabilities])
            probD[str(counter).zfill(4)] = content
            counter += 1
        return probD

    #############################
    ### Miscellaneous Methods ###
    #############################

    def _laplacians(self, paths):
        # decode on the thread pool (cv2.imread releases the GIL; at 800x600 the PNG decode is most of processRaw)
        imgs = self._per_frame(lambda pth: cv2.imread(pth, 0), list(paths))
        if not imgs:
            return []
        return [np.float64(v) for v in self._context().laplacian_var(imgs)]

    def Laplacian(self, imagePath):
        ''' blurriness factor: variance of the Laplacian (reference analyze.py:441-445); decode stays on cv2 '''
        img = cv2.imread(imagePath, 0)
        return np.float64(self._context().laplacian_var([img])[0])   # ndarray.var() returns np.float64
