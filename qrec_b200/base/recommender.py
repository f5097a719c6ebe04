"""`Recommender`: the plugin life cycle every model runs through.

Mirrors base/recommender.py:14-212 of the reference: the constructor signature, the attribute
names other code reads, and the fixed order of `execute()`:
    readConfiguration -> initializing_log -> printAlgorConfig (fold '[1]') -> initModel ->
    trainModel (or trainModel_tf when `-tf` is set and TensorFlow imports) -> evalRanking |
    evalRatings -> saveModel.
Evaluation keeps the reference's observable behaviour (rated items are scored 0 rather than
removed, -topN is clamped to <= 100, result files and their names), but asks the model for
scores in blocks (`score_block`) when the model offers it.
"""
import sys
from os.path import abspath
from time import strftime, localtime, time

from ..data.rating import Rating
from ..util.config import OptionConf
from ..util.io import FileIO
from ..util.log import Log
from ..util.measure import Measure
from ..util.qmath import find_k_largest


def _now():
    return strftime("%Y-%m-%d %H-%M-%S", localtime(time()))


class Recommender(object):
    def __init__(self, conf, trainingSet, testSet, fold='[1]'):
        self.config = conf
        self.isSaveModel = False
        self.isLoadModel = False
        self.isOutput = True
        self.ranking = None
        self.output = None
        self.data = Rating(self.config, trainingSet, testSet)
        self.foldInfo = fold
        self.evalSettings = OptionConf(self.config['evaluation.setup'])
        self.measure = []
        self.recOutput = []
        self.num_users, self.num_items, self.train_size = self.data.trainingSize()

    # ------------------------------------------------------------------ life-cycle hooks
    def readConfiguration(self):
        self.modelName = self.config['model.name']
        self.output = OptionConf(self.config['output.setup'])
        self.isOutput = self.output.isMainOn()
        self.ranking = OptionConf(self.config['item.ranking'])

    def initializing_log(self):
        self.log = Log(self.modelName, self.modelName + self.foldInfo + ' ' + _now())
        self.log.add('### model configuration ###')
        for key in self.config.config:
            self.log.add(key + '=' + self.config[key])

    def printAlgorConfig(self):
        print('Model:', self.config['model.name'])
        print('Ratings dataset:', abspath(self.config['ratings']))
        if self.evalSettings.contains('-testSet'):
            print('Test set:', abspath(self.evalSettings['-testSet']))
        print('Training set size: (user count: %d, item count %d, record count: %d)' % self.data.trainingSize())
        print('Test set size: (user count: %d, item count %d, record count: %d)' % self.data.testSize())
        print('=' * 80)
        name = self.config['model.name']
        if self.config.contains(name):
            args = OptionConf(self.config[name])
            print('Specific parameters:', ''.join(k[1:] + ':' + args[k] + '  ' for k in args.keys()))
            print('=' * 80)

    def initModel(self):
        pass

    def trainModel(self):
        pass

    buildModel = trainModel          # BASELINE.json's name for the same hook

    def trainModel_tf(self):
        pass

    def saveModel(self):
        pass

    def loadModel(self):
        pass

    def predictForRating(self, u, i):
        pass

    def predictForRanking(self, u):
        pass

    # ------------------------------------------------------------------ evaluation
    def checkRatingBoundary(self, prediction):
        lo, hi = self.data.rScale[0], self.data.rScale[-1]
        if prediction > hi:
            return hi
        if prediction < lo:
            return lo
        return round(prediction, 3)

    def evalRatings(self):
        lines = ['userId  itemId  original  prediction\n']
        for pos, (user, item, rating) in enumerate(self.data.testData):
            pred = self.checkRatingBoundary(self.predictForRating(user, item))
            self.data.testData[pos].append(pred)
            lines.append(user + ' ' + item + ' ' + str(rating) + ' ' + str(pred) + '\n')
        stamp = _now()
        out_dir = self.output['-dir']
        if self.isOutput:
            FileIO.writeFile(out_dir, self.config['model.name'] + '@' + stamp + '-rating-predictions' + self.foldInfo + '.txt', lines)
            print('The result has been output to ', abspath(out_dir), '.')
        self.measure = Measure.ratingMeasure(self.data.testData)
        FileIO.writeFile(out_dir, self.config['model.name'] + '@' + stamp + '-measure' + self.foldInfo + '.txt', self.measure)
        self.log.add('###Evaluation Results###')
        self.log.add(self.measure)
        print('The result of %s %s:\n%s' % (self.modelName, self.foldInfo, ''.join(self.measure)))

    def _top_n_setting(self):
        if not self.ranking.contains('-topN'):
            print('No correct evaluation metric is specified!')
            sys.exit(-1)
        top = [int(x) for x in self.ranking['-topN'].split(',')]
        N = max(top)
        if N > 100 or N < 1:
            print('N can not be larger than 100! It has been reassigned to 10')
            N = 10
        return top, N

    def _recommend(self, user, N):
        """Top-N (item name, score) list of one user; rated items are overwritten with score 0
        (base/recommender.py:147-149), not excluded."""
        scores = self.predictForRanking(user)
        for item in self.data.userRated(user)[0]:
            scores[self.data.item[item]] = 0
        ids, vals = find_k_largest(N, scores)
        return [(self.data.id2item[k], v) for k, v in zip(ids, vals)]

    def device_tables(self):
        """(U, V) fp32 CUDA tensors such that score(u, i) = U[u] . V[i], or None if the model ranks
        some other way (then evaluation stays on the host, one user at a time)."""
        return None

    def _recommend_all_on_device(self, N):
        """`engine=... -eval gpu`: top-N lists of every known test user in a few large launches."""
        if not self.config.contains('engine') or OptionConf(self.config['engine']).options.get('-eval') != 'gpu':
            return None
        tables = self.device_tables()
        if tables is None:
            return None
        from ..evaluate import batched_top_n
        users = [u for u in self.data.testSet_u if self.data.containsUser(u)]
        ids, vals = batched_top_n(tables[0], tables[1], [self.data.user[u] for u in users], self.data.rated_csr(), N)
        return {u: [(self.data.id2item[int(k)], float(v)) for k, v in zip(ids[r], vals[r])] for r, u in enumerate(users)}

    def evalRanking(self):
        top, N = self._top_n_setting()
        self.recOutput.append('userId: recommendations in (itemId, ranking score) pairs, * means the item matches.\n')
        recList = {}
        n_test = len(self.data.testSet_u)
        batched = self._recommend_all_on_device(N)
        for pos, user in enumerate(self.data.testSet_u):
            recList[user] = batched[user] if batched is not None and user in batched else self._recommend(user, N)
            if pos % 100 == 0:
                print(self.modelName, self.foldInfo, 'progress:' + str(pos) + '/' + str(n_test))
            truth = self.data.testSet_u[user]
            self.recOutput.append(user + ':' + ''.join(
                ' (' + name + ',' + str(score) + ')' + ('*' if name in truth else '') for name, score in recList[user]) + '\n')
        stamp = _now()
        out_dir = self.output['-dir']
        if self.isOutput:
            FileIO.writeFile(out_dir, self.config['model.name'] + '@' + stamp + '-top-' + str(N) + 'items' + self.foldInfo + '.txt', self.recOutput)
            print('The result has been output to ', abspath(out_dir), '.')
        if self.evalSettings.contains('-predict'):
            sys.exit(0)
        self.measure = Measure.rankingMeasure(self.data.testSet_u, recList, top)
        self.log.add('###Evaluation Results###')
        self.log.add(self.measure)
        FileIO.writeFile(out_dir, self.config['model.name'] + '@' + stamp + '-measure' + self.foldInfo + '.txt', self.measure)
        print('The result of %s %s:\n%s' % (self.modelName, self.foldInfo, ''.join(self.measure)))

    # ------------------------------------------------------------------ driver
    def execute(self):
        self.readConfiguration()
        self.initializing_log()
        if self.foldInfo == '[1]':
            self.printAlgorConfig()
        if self.isLoadModel:
            print('Loading model %s...' % self.foldInfo)
            self.loadModel()
        else:
            print('Initializing model %s...' % self.foldInfo)
            self.initModel()
            print('Building Model %s...' % self.foldInfo)
            # `-tf` selects the minibatch/Adam variant (base/recommender.py:195-203 of the reference).  The
            # engine never imports TensorFlow, so the switch does not depend on it being installed; a model
            # that does not define its own trainModel_tf (the reference's base hook is `pass`, which would
            # evaluate the random initial tables) falls back to trainModel with a warning.
            own_tf = type(self).trainModel_tf is not Recommender.trainModel_tf
            if self.evalSettings.contains('-tf') and own_tf:
                self.trainModel_tf()
            else:
                if self.evalSettings.contains('-tf'):
                    print('WARNING: %s has no trainModel_tf; `-tf` ignored, running trainModel().' % self.modelName)
                self.trainModel()
        print('Predicting %s...' % self.foldInfo)
        if self.ranking.isMainOn():
            self.evalRanking()
        else:
            self.evalRatings()
        if self.isSaveModel:
            print('Saving model %s...' % self.foldInfo)
            self.saveModel()
        return self.measure
