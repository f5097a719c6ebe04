"""Ranking / rating metrics, same definitions and output strings as util/measure.py:24-138
(precision = hits / (users*N); recall averaged per user; NDCG with 1/ln(rank+2) and IDCG over
min(N, |test items|))."""
import math


class Measure(object):
    @staticmethod
    def hits(origin, res):
        return {user: len(set(origin[user]).intersection(item[0] for item in res[user])) for user in origin}

    @staticmethod
    def precision(hits, N):
        return sum(hits.values()) / (len(hits) * N)

    @staticmethod
    def recall(hits, origin):
        per_user = [hits[user] / len(origin[user]) for user in hits]
        return sum(per_user) / len(per_user)

    @staticmethod
    def F1(prec, recall):
        return 2 * prec * recall / (prec + recall) if (prec + recall) != 0 else 0

    @staticmethod
    def NDCG(origin, res, N):
        # plain left-to-right accumulation, like the reference: Python >= 3.12's sum() compensates
        # (Neumaier) and would change the last digit of the printed metric
        total = 0
        for user in res:
            dcg = 0
            for rank, item in enumerate(res[user]):
                if item[0] in origin[user]:
                    dcg += 1.0 / math.log(rank + 2)
            idcg = 0
            for rank in range(min(N, len(origin[user]))):
                idcg += 1.0 / math.log(rank + 2)
            total += dcg / idcg
        return total / len(res)

    @staticmethod
    def rankingMeasure(origin, res, N):
        out = []
        for n in N:
            predicted = {user: res[user][:n] for user in res}
            if len(origin) != len(predicted):
                print('The Lengths of test set and predicted set are not match!')
                raise SystemExit(-1)
            hits = Measure.hits(origin, predicted)
            prec = Measure.precision(hits, n)
            recall = Measure.recall(hits, origin)
            out.append('Top ' + str(n) + '\n')
            out.append('Precision:' + str(prec) + '\n')
            out.append('Recall:' + str(recall) + '\n')
            out.append('F1:' + str(Measure.F1(prec, recall)) + '\n')
            out.append('NDCG:' + str(Measure.NDCG(origin, predicted, n)) + '\n')
        return out

    @staticmethod
    def MAE(res):
        error = 0
        for e in res:
            error += abs(e[2] - e[3])
        return error / len(res) if res else error

    @staticmethod
    def RMSE(res):
        error = 0
        for e in res:
            error += (e[2] - e[3]) ** 2
        return math.sqrt(error / len(res)) if res else error

    @staticmethod
    def ratingMeasure(res):
        return ['MAE:' + str(Measure.MAE(res)) + '\n', 'RMSE:' + str(Measure.RMSE(res)) + '\n']
