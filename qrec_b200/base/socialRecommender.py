"""`SocialRecommender`: an IterativeRecommender with a cleaned `Social` view (reference:
base/socialRecommender.py:5-54): relations whose endpoints are not training users are dropped from the
followee / follower dictionaries and from the relation list."""
from os.path import abspath

from ..data.social import Social
from ..util import config
from .iterativeRecommender import IterativeRecommender


class SocialRecommender(IterativeRecommender):
    def __init__(self, conf, trainingSet, testSet, relation, fold='[1]'):
        super(SocialRecommender, self).__init__(conf, trainingSet, testSet, fold)
        self.social = Social(self.config, relation)
        known = self.data.user
        for table in (self.social.followees, self.social.followers):
            for u in [u for u in table if u not in known]:
                del table[u]
            for u in table:
                for v in [v for v in table[u] if v not in known]:
                    del table[u][v]
        self.social.relation[:] = [r for r in self.social.relation if r[0] in known and r[1] in known]

    def readConfiguration(self):
        super(SocialRecommender, self).readConfiguration()
        self.regS = float(config.OptionConf(self.config['reg.lambda'])['-s'])

    def printAlgorConfig(self):
        super(SocialRecommender, self).printAlgorConfig()
        print('Social dataset:', abspath(self.config['social']))
        print('Social relation size ', '(User count:', len(self.social.user), 'Relation count:' + str(len(self.social.relation)) + ')')
        print('Social Regularization parameter: regS %.3f' % (self.regS))
        print('=' * 80)
