"""Partitioning of the interaction list into training and test parts.

Behavioural contract (reference: util/dataSplit.py:9-44): the hold-out split spends exactly one
`random()` draw of Python's global MT19937 per record, in record order, so a seeded run leaves
the generator where the reference would and the samplers that follow see the same stream; the
k-fold generator assigns record p to fold p mod k.  With binarised data, records whose rating
is falsy never enter a test part.
"""
import random as _random

from .io import FileIO


class DataSplit(object):
    @staticmethod
    def dataSplit(data, test_ratio=0.3, output=False, path='./', order=1, binarized=False):
        ratio = test_ratio if 0 < test_ratio < 1 else 0.3
        draw = _random.random
        to_test = [draw() < ratio for _ in data]            # one draw per record, in order
        trainingSet = [rec for rec, t in zip(data, to_test) if not t]
        testSet = [rec for rec, t in zip(data, to_test) if t and (rec[2] or not binarized)]
        if output:
            FileIO.writeFile(path, 'testSet[%s]' % order, testSet)
            FileIO.writeFile(path, 'trainingSet[%s]' % order, trainingSet)
        return trainingSet, testSet

    @staticmethod
    def crossValidation(data, k, output=False, path='./', order=1, binarized=False):
        folds = k if 1 < k <= 10 else 3
        for held_out in range(folds):
            train_part = [list(rec) for pos, rec in enumerate(data) if pos % folds != held_out]
            test_part = [list(rec) for pos, rec in enumerate(data)
                         if pos % folds == held_out and (rec[2] or not binarized)]
            yield train_part, test_part
