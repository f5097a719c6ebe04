"""Array-backed interaction tables and their binary cache.

The reference keeps a dataset as a Python list of [user, item, rating] records and two dict-of-dicts
(util/io.py:31-76, data/rating.py:33-67): fine for FilmTrust, ~20 GB and minutes for 50 M lines.
`InteractionTable` holds the same information as flat arrays -- dense ids assigned in order of FIRST
APPEARANCE exactly like Rating.__generateSet, so the id space (and therefore every sampled index and
every table row) is the one the reference would produce -- and round-trips through one .npz file.
"""
import os
import re

import numpy as np

DEFAULT_DELIM = ' |,|\\t'          # util/io.py:36: the reference's default `-delim`


class InteractionTable(object):
    def __init__(self, user_names, item_names, u, i, r):
        self.user_names = np.asarray(user_names)
        self.item_names = np.asarray(item_names)
        self.u = np.ascontiguousarray(u, dtype=np.int32)
        self.i = np.ascontiguousarray(i, dtype=np.int32)
        self.r = np.ascontiguousarray(r, dtype=np.float64)
        assert self.u.shape == self.i.shape == self.r.shape

    # ------------------------------------------------------------------ construction
    @staticmethod
    def _first_appearance_ids(names):
        """names -> (ids, vocabulary in first-appearance order); what `if name not in d: d[name]=len(d)` does."""
        names = np.asarray(names)
        vocab, first, inv = np.unique(names, return_index=True, return_inverse=True)
        order = np.argsort(first, kind='stable')              # vocabulary sorted by first appearance
        rank = np.empty(len(vocab), dtype=np.int64)
        rank[order] = np.arange(len(vocab))
        return rank[inv].astype(np.int32), vocab[order]

    @classmethod
    def from_records(cls, records):
        """records: iterable of [user, item, rating] (the reference's in-memory format)."""
        users = [rec[0] for rec in records]
        items = [rec[1] for rec in records]
        r = np.array([rec[2] for rec in records], dtype=np.float64)
        u, un = cls._first_appearance_ids(np.array(users, dtype=object).astype(str)) if users else (np.zeros(0, np.int32), np.zeros(0, str))
        i, inn = cls._first_appearance_ids(np.array(items, dtype=object).astype(str)) if items else (np.zeros(0, np.int32), np.zeros(0, str))
        return cls(un, inn, u, i, r)

    @classmethod
    def _from_text_native(cls, path, columns, header, binarize_threshold):
        """qrec_text_load (csrc/host_loader.cpp): one pass over the file in C++, hash-map id assignment.
        Returns None when the native reader declines the file (short line, unusual number syntax): the
        Python loop below then decides, with the reference's own behaviour."""
        import ctypes as C
        from .._lib import lib
        col_r = columns[2] if len(columns) >= 3 else -1
        if binarize_threshold is not None and col_r < 0:
            return None
        h = lib.qrec_text_load(os.fsencode(path), int(columns[0]), int(columns[1]), int(col_r), int(bool(header)),
                               int(binarize_threshold is not None), float(binarize_threshold or 0.0))
        if not h:
            if not os.path.exists(path):
                raise FileNotFoundError(path)
            return None
        try:
            n = lib.qrec_text_rows(h)
            u, i, r = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.float64)
            lib.qrec_text_copy(h, u.ctypes.data_as(C.POINTER(C.c_int32)), i.ctypes.data_as(C.POINTER(C.c_int32)),
                               r.ctypes.data_as(C.POINTER(C.c_double)))
            names = []
            for which in (0, 1):
                size = lib.qrec_text_names(h, which, None, 0)
                buf = C.create_string_buffer(max(1, size))
                lib.qrec_text_names(h, which, buf, size)
                count = lib.qrec_text_vocab_size(h, which)
                names.append(np.array(buf.raw[:size].decode().split('\n')) if count else np.zeros(0, str))
        except UnicodeDecodeError:
            return None
        finally:
            lib.qrec_text_free(h)
        return cls(names[0], names[1], u, i, r)

    @classmethod
    def from_text(cls, path, columns=(0, 1, 2), delim=DEFAULT_DELIM, header=False, binarize_threshold=None):
        """Same parsing rules as FileIO.loadDataSet (util/io.py:31-76): regex-split lines, `columns`
        picks user / item / rating, rows with rating < threshold are dropped and the rest set to 1 when
        binarising, a missing rating column means 1."""
        if delim == DEFAULT_DELIM:
            table = cls._from_text_native(path, columns, header, binarize_threshold)
            if table is not None:
                return table
        splitter = re.compile(delim)
        users, items, ratings = [], [], []
        with open(path) as fh:
            for lineno, line in enumerate(fh):
                if header and lineno == 0:
                    continue
                f = splitter.split(line.strip())
                rating = float(f[columns[2]]) if len(columns) >= 3 else 1.0
                if binarize_threshold is not None:
                    if float(f[columns[2]]) < binarize_threshold:
                        continue
                    rating = 1.0
                users.append(f[columns[0]]); items.append(f[columns[1]]); ratings.append(rating)
        u, un = cls._first_appearance_ids(np.array(users)) if users else (np.zeros(0, np.int32), np.zeros(0, str))
        i, inn = cls._first_appearance_ids(np.array(items)) if items else (np.zeros(0, np.int32), np.zeros(0, str))
        return cls(un, inn, u, i, np.array(ratings, dtype=np.float64))

    # ------------------------------------------------------------------ cache
    def save(self, path):
        np.savez_compressed(path, user_names=self.user_names.astype(str), item_names=self.item_names.astype(str),
                            u=self.u, i=self.i, r=self.r)

    @classmethod
    def load(cls, path):
        z = np.load(path if path.endswith('.npz') else path + '.npz')
        return cls(z['user_names'], z['item_names'], z['u'], z['i'], z['r'])

    @classmethod
    def cached_from_text(cls, path, cache_dir=None, **kw):
        """Parses `path` once; later calls load the .npz next to it (or in cache_dir) if it is newer."""
        key = re.sub(r'[^A-Za-z0-9]+', '_', repr(sorted(kw.items())))
        base = os.path.join(cache_dir or os.path.dirname(os.path.abspath(path)),
                            '.' + os.path.basename(path) + '.' + key + '.qrec.npz')
        if os.path.exists(base) and os.path.getmtime(base) >= os.path.getmtime(path):
            return cls.load(base)
        table = cls.from_text(path, **kw)
        try:
            table.save(base)
        except OSError:
            pass                                   # read-only dataset directory: just skip the cache
        return table

    # ------------------------------------------------------------------ views
    @property
    def num_users(self):
        return len(self.user_names)

    @property
    def num_items(self):
        return len(self.item_names)

    def __len__(self):
        return self.u.shape[0]

    def to_records(self):
        """Back to the reference's list-of-records format (for the classic constructors)."""
        un, inn = self.user_names.tolist(), self.item_names.tolist()
        return [[un[a], inn[b], float(c)] for a, b, c in zip(self.u.tolist(), self.i.tolist(), self.r.tolist())]

    def rated_csr(self, positive_threshold=1.0):
        from ..engine import RatedCSR
        return RatedCSR(self.num_users, self.num_items, self.u, self.i, self.r, positive_threshold)

    def split(self, keep_mask):
        """Train/test split by a boolean mask over the records (e.g. MT19937.data_split); the training part
        gets fresh first-appearance ids, like Rating built from the training list alone."""
        keep_mask = np.asarray(keep_mask, dtype=bool)
        un, inn = self.user_names, self.item_names
        tr = InteractionTable.from_records(list(zip(un[self.u[keep_mask]], inn[self.i[keep_mask]], self.r[keep_mask])))
        te = (un[self.u[~keep_mask]], inn[self.i[~keep_mask]], self.r[~keep_mask])
        return tr, te
