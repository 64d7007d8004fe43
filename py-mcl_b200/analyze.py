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
import time

import cv2
import numpy as np

from . import _lib
from .engine import ShardedMap
from .Matcher import Matcher, angle_window

extension = '.png'


class analyzer(object):

    def __init__(self, method, width, height, numLocations=7):
        self.numLocations = numLocations
        self.indices = [None] * self.numLocations
        self.method = method
        self.w = width
        self.h = height
        self.rawP = []
        self.blurP = []
        self.commands = self.readCommand('commands.txt')
        self.bestGuess = []
        self.orb_mode = 'knn'
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
        matcher = Matcher(self.method, width=self.w, height=self.h)
        if self.method != 'BOW':
            for i in range(self.numLocations):
                matcher.setDirectory('map/' + str(i))
                if self.method != 'Color':
                    self.indices[i] = matcher.createFeatureIndex()
                else:
                    raise NotImplementedError("colour-histogram matching is outside the accelerated hot path")
        else:
            matcher.writeIndices()
        self._map = None

    ####################
    ### Main Methods ###
    ####################

    def _kind(self):
        return self.method if self.method in ('SIFT', 'SURF') else 'ORB'

    def _mode(self):
        return _lib.CROSSCHECK if (self._kind() == 'ORB' and self.orb_mode == 'crosscheck') else _lib.KNN2_RATIO

    def _angle_paths(self, loc):
        return ['map/' + str(loc) + '/angle' + str(i).zfill(3) + extension for i in range(0, 375, 15)]

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
        import os
        from concurrent.futures import ThreadPoolExecutor
        kind = kind or self._kind()

        def one(path):
            m = Matcher(self.method, width=self.w, height=self.h)
            m.setQuery(path)
            return m._query_descriptors(kind)

        def many(paths):
            workers = max(1, min(len(paths), int(os.environ.get("MCL_EXTRACT_THREADS", os.cpu_count() or 1))))
            if workers <= 1:
                return [one(p) for p in paths]
            with ThreadPoolExecutor(max_workers=workers) as pool:
                return list(pool.map(one, paths))

        # one process per GPU: the frames are dealt to the ranks, each extracts its share, descriptors are all-gathered
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
        matcher = Matcher(self.method, width=self.w, height=self.h)
        print('Matching...')
        frames = glob.glob('cam1_img' + '/*' + extension)

        if self.method != 'BOW':
            qdes = self._extract_many(frames)
            counts = self._resident_map().match_counts(qdes, mode=self._mode())
            for f, imagePath in enumerate(frames):
                for i in range(self.numLocations):
                    p.append(self._run_result(counts[f, i * 25:(i + 1) * 25]))
                print('\t' + imagePath)
        else:
            import joblib
            locs = []
            for i in range(self.numLocations):
                im_features, image_paths, idf, numWords, voc = joblib.load('map/' + str(i) + '.pkl')
                locs.append((im_features, idf, voc))
            bow = _lib.BowIndex(self._context(), locs, int(np.asarray(locs[0][0]).shape[1]))
            qdes = self._extract_many(frames, 'SIFT')
            scores, _ = bow.score(qdes)
            for f, imagePath in enumerate(frames):
                for i in range(self.numLocations):
                    s = scores[f:f + 1, bow.img_off[i]:bow.img_off[i + 1]]
                    p.append([10 * np.max(s), s[0].tolist()])
                print('\t' + imagePath)
            bow.close()
        self.rawP = p
        self.writeProb(p, 'rawP.txt', 'w')

        end = time.time()
        print('Time elapsed: %0.1f' % (end - start))

    def _filter(self, command, previousProbs, p, blurFactor):
        """accountCommand -> prevWeight -> probUpdate -> best guess in one kernel (mcl_filter_step).
        Reproduces the reference's aliasing: the command shift is written back into the SAME inner lists of
        previousProbs (analyze.py:317-335 mutate a shallow copy)."""
        prev = np.array([[c[0]] + list(c[1]) for c in previousProbs], dtype=np.float64)
        cur = np.array([[c[0]] + list(c[1]) for c in p], dtype=np.float64)
        prev2, out, best = self._context().filter_step(_lib.F_ALL, command, float(blurFactor), prev, cur)
        if command == 'l' or command == 'r':
            for circles, row in zip(previousProbs, prev2):
                circles[1] = row[1:].tolist()
        elif command == 'f':
            for circles, row, old in zip(previousProbs, prev2, prev):
                if row[0] != old[0]:
                    circles[0] = float(row[0])
        adjusted = [[float(row[0]), row[1:].tolist()] for row in out]
        return adjusted, best

    def processRaw(self):
        """
        Processes the raw data from rawP.txt (reference analyze.py:98-146).
        """
        previousProbs = []
        for i in range(self.numLocations):
            previousProbs.append([1, [1/75] * 25])

        start = time.time()
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
        this is one masked GPU call per frame.
        """
        if self.method != 'BOW':
            print('Creating indices...')
            self.createIndex()

        blurP = []
        previousProbs = []
        bestAngleIndex = None
        bestCircleIndex = None

        for i in range(self.numLocations):
            previousProbs.append([1, [1/75] * 25])

        matcher = Matcher(self.method, width=self.w, height=self.h)
        start = time.time()
        print('Matching...')
        resident = self._resident_map()
        n_img = self.numLocations * 25

        for imagePath in glob.glob('cam1_img' + '/*' + extension):
            des = self._extract(matcher, imagePath)
            window = angle_window(bestAngleIndex) if bestAngleIndex is not None else [True] * 25
            if bestCircleIndex is None:
                locs = list(range(self.numLocations))
            else:
                lower = bestCircleIndex - 2
                upper = bestCircleIndex + 2
                locs = [i for i in range(self.numLocations) if i >= lower and i <= upper]
            mask = np.zeros((1, n_img), dtype=np.uint8)
            for i in locs:
                mask[0, i * 25:(i + 1) * 25] = window
            counts = resident.match_counts([des], mode=self._mode(), mask=mask, fill_value=1)[0]
            p = []
            for i in range(self.numLocations):
                if i in locs:
                    p.append(self._run_result(counts[i * 25:(i + 1) * 25]))
                else:
                    p.append([1, [1/75] * 25])
            print('\t' + imagePath)

            command = self.commands[imagePath.replace('cam1_img/', '').replace(extension, '')]
            blurFactor = self.Laplacian(imagePath)
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

    def _stage(self, stages, previousP, currentP, blurFactor=0.0, command='s'):
        prev = np.array([[c[0]] + list(c[1]) for c in previousP], dtype=np.float64)
        cur = np.array([[c[0]] + list(c[1]) for c in currentP], dtype=np.float64)
        return self._context().filter_step(stages, command, float(blurFactor), prev, cur)

    def probUpdate(self, previousP, currentP, blurFactor):
        """Blur-proportional blend (reference analyze.py:239-276)."""
        _, out, _ = self._stage(_lib.F_BLUR, previousP, currentP, blurFactor)
        return [[float(r[0]), r[1:].tolist()] for r in out]

    def prevWeight(self, previousP, currentP):
        """0.7 / 0.3 blend (reference analyze.py:278-310)."""
        _, out, _ = self._stage(_lib.F_PREV, previousP, currentP)
        return [[float(r[0]), r[1:].tolist()] for r in out]

    def accountCommand(self, command, previousP):
        """Shift particles according to the current command (reference analyze.py:312-336); mutates the inner
        lists of previousP like the reference's shallow copy does."""
        prev = np.array([[c[0]] + list(c[1]) for c in previousP], dtype=np.float64)
        prev2, _, _ = self._context().filter_step(_lib.F_COMMAND, command, 0.0, prev, prev)
        copy = previousP[:]
        if command == 'l' or command == 'r':
            for circles, row in zip(copy, prev2):
                circles[1] = row[1:].tolist()
        elif command == 'f':
            for circles, row, old in zip(copy, prev2, prev):
                if row[0] != old[0]:
                    circles[0] = float(row[0])
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
        for index in prob:
            file.write(str(index[0]) + '\n')
            file.write(str(index[1]) + '\n')
        file.close()

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
        imgs = [cv2.imread(pth, 0) for pth in paths]
        if not imgs:
            return []
        return [float(v) for v in self._context().laplacian_var(imgs)]

    def Laplacian(self, imagePath):
        ''' blurriness factor: variance of the Laplacian (reference analyze.py:441-445); decode stays on cv2 '''
        img = cv2.imread(imagePath, 0)
        return float(self._context().laplacian_var([img])[0])
