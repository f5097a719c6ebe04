"""BPR-MF at dataset sizes the record-list API cannot hold: trains straight from an
`InteractionTable` (flat id arrays, the reference's id space) with the fused user-major epoch kernel --
device-side Philox negatives, P[u] sequential inside a user -- and keeps the reference's epoch
bookkeeping: loss = sum -ln s + regU|P|^2 + regI|Q|^2 (model/ranking/BPR.py:40,53), the adaptive learning
rate and the |delta loss| < 1e-3 stop of base/iterativeRecommender.py:56-63,82-102.
"""
import numpy as np


class ScaleBPR(object):
    def __init__(self, table, emb_size=64, lr=0.01, max_lr=1.0, reg_u=0.001, reg_i=0.001, device=0, seed=0):
        import torch
        self.table = table
        self.emb_size, self.lRate, self.maxLRate, self.regU, self.regI = emb_size, lr, max_lr, reg_u, reg_i
        self.device = self._make_device(device)
        self.seed = seed
        csr = table.rated_csr()
        self.num_users, self.num_items = table.num_users, table.num_items
        dev = self.device
        # positives in the reference's iteration order (CSR, insertion order inside a user)
        self.rowptr = torch.from_numpy(csr.pos_rowptr).to(dev)
        self.pos_items = torch.from_numpy(csr.pos_cols).to(dev)
        self.rated_rowptr = torch.from_numpy(csr.sorted_rowptr).to(dev)
        self.rated_cols = torch.from_numpy(csr.sorted_cols).to(dev)
        self.csr = csr
        # P = rand/3, Q = rand/3 from numpy's legacy stream, P first (base/iterativeRecommender.py:37-38)
        self.dpad = emb_size + (-emb_size) % 4
        P = np.zeros((self.num_users, self.dpad), np.float32)
        Q = np.zeros((self.num_items, self.dpad), np.float32)
        P[:, :emb_size] = np.random.rand(self.num_users, emb_size) / 3
        Q[:, :emb_size] = np.random.rand(self.num_items, emb_size) / 3
        self.P, self.Q = torch.from_numpy(P).to(dev), torch.from_numpy(Q).to(dev)
        self._acc = torch.zeros(3, dtype=torch.float64, device=dev)
        self.loss, self.lastLoss, self.epoch = 0.0, 0.0, 0
        self.history = []

    @staticmethod
    def _make_device(index):
        """The CUDA device (made current); the one place this class touches torch.cuda."""
        import torch
        dev = torch.device('cuda', index)
        torch.cuda.set_device(dev)
        return dev

    def run_epoch(self):
        """One epoch; returns True when the reference's convergence test fires."""
        from . import engine as E
        self._acc.zero_()
        E.bpr_epoch_usermajor(self.P, self.Q, self.rowptr, self.pos_items, self.rated_rowptr, self.rated_cols,
                              self.num_items, self.seed, self.epoch, self.lRate, self.regU, self.regI, self._acc[0:1])
        E.sumsq(self.P, self._acc[1:2])
        E.sumsq(self.Q, self._acc[2:3])
        a = self._acc.cpu().numpy()
        self.loss = float(a[0] + self.regU * a[1] + self.regI * a[2])
        self.epoch += 1
        if np.isnan(self.loss):
            raise FloatingPointError('Loss = NaN or Infinity: current settings does not fit the recommender!')
        delta = self.lastLoss - self.loss
        self.history.append((self.epoch, self.loss, delta, self.lRate))
        converged = abs(delta) < 1e-3
        if not converged:
            if self.epoch > 1:
                self.lRate *= 1.05 if abs(self.lastLoss) > abs(self.loss) else 0.5
            if self.lRate > self.maxLRate > 0:
                self.lRate = self.maxLRate
        self.lastLoss = self.loss
        return converged

    def fit(self, max_epoch):
        for _ in range(max_epoch):
            if self.run_epoch():
                break
        return self

    def tables(self):
        """(P, Q) as host numpy arrays of the logical width."""
        d = self.emb_size
        return self.P[:, :d].cpu().numpy(), self.Q[:, :d].cpu().numpy()

    def top_n(self, user_ids, N=10, block=2048):
        """Top-N item ids and scores per user (rated items scored 0, as base/recommender.py:147-149)."""
        from .evaluate import batched_top_n
        return batched_top_n(self.P, self.Q, user_ids, self.csr, N, block=block)

    def evaluate(self, test_table_or_triples, tops=(10,), block=2048):
        """Ranking metrics (the reference's definitions, util/measure.py:24-138) for the test interactions of
        users known from training.  `test_table_or_triples`: (user_names, item_names, ratings) arrays or a
        list of [user, item, rating] records.  Returns the reference-format list of metric strings."""
        from .util.fastmeasure import ranking_measures
        if isinstance(test_table_or_triples, tuple):
            tu, ti = (np.asarray(x) for x in test_table_or_triples[:2])
        else:
            tu = np.array([rec[0] for rec in test_table_or_triples])
            ti = np.array([rec[1] for rec in test_table_or_triples])
        lut_u = {n: k for k, n in enumerate(self.table.user_names.tolist())}
        lut_i = {n: k for k, n in enumerate(self.table.item_names.tolist())}
        # the reference keeps test users unknown to training (they get globalMean scores); this path ranks
        # the known ones -- items unseen in training can never be recommended and only count in |test_u|
        next_unknown = len(lut_i)
        pairs = {}
        for a, b in zip(tu.tolist(), ti.tolist()):
            if a in lut_u:
                if b not in lut_i:
                    lut_i[b] = next_unknown
                    next_unknown += 1
                pairs.setdefault(lut_u[a], {})[lut_i[b]] = 1
        users = list(pairs)
        rowptr = np.zeros(len(users) + 1, np.int64)
        rowptr[1:] = np.cumsum([len(pairs[u]) for u in users])
        cols = np.array([c for u in users for c in pairs[u]], dtype=np.int64)
        N = max(tops)
        ids, _ = self.top_n(users, N, block=block)
        return ranking_measures(ids, rowptr, cols, list(tops))
