"""Train/test partitioning (reference: util/dataSplit.py:9-44).  The Bernoulli split draws one
random() per record from Python's global MT19937, exactly like the reference, so a seeded run
leaves the generator in the same state for the samplers that follow."""
from random import random

from .io import FileIO


class DataSplit(object):
    @staticmethod
    def dataSplit(data, test_ratio=0.3, output=False, path='./', order=1, binarized=False):
        if not 0 < test_ratio < 1:
            test_ratio = 0.3
        train, test = [], []
        for rec in data:
            if random() < test_ratio:
                if not binarized or rec[2]:
                    test.append(rec)
            else:
                train.append(rec)
        if output:
            FileIO.writeFile(path, 'testSet[' + str(order) + ']', test)
            FileIO.writeFile(path, 'trainingSet[' + str(order) + ']', train)
        return train, test

    @staticmethod
    def crossValidation(data, k, output=False, path='./', order=1, binarized=False):
        if k <= 1 or k > 10:
            k = 3
        for fold in range(k):
            train, test = [], []
            for pos, rec in enumerate(data):
                if pos % k == fold:
                    if not binarized or rec[2]:
                        test.append(rec[:])
                else:
                    train.append(rec[:])
            yield train, test
