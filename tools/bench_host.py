#!/usr/bin/env python
"""Host-side pieces of the data path, timed on the CPU (no GPU needed): the native rating-file reader
against the Python loop, and the native per-user item-set builder against the numpy construction it
replaced.  One JSON line each.   python tools/bench_host.py [--lines 5000000] [--pairs 20000000]"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--lines', type=int, default=5_000_000)
    ap.add_argument('--pairs', type=int, default=20_000_000)
    args = ap.parse_args()
    from qrec_b200 import engine as E
    from qrec_b200.data.interactions import InteractionTable
    from oracle import bpr_oracle as O                      # the numpy construction lives with the checkers
    cores = os.cpu_count()
    rng = np.random.default_rng(0)
    n = args.lines
    u, i, r = rng.integers(0, n // 25, n), rng.integers(0, n // 100, n), rng.integers(1, 6, n)
    path = os.path.join(tempfile.mkdtemp(), 'ratings.txt')
    with open(path, 'w') as f:
        f.write('\n'.join('u%d i%d %d' % t for t in zip(u.tolist(), i.tolist(), r.tolist())) + '\n')
    t0 = time.perf_counter(); a = InteractionTable.from_text(path); t_native = time.perf_counter() - t0
    t0 = time.perf_counter(); b = InteractionTable.from_text(path, delim='[ ,\t]'); t_python = time.perf_counter() - t0
    same = (a.user_names.tolist() == b.user_names.tolist() and np.array_equal(a.u, b.u) and np.array_equal(a.i, b.i)
            and np.array_equal(a.r, b.r))
    print(json.dumps({'what': 'rating-file reader', 'lines': n, 'file_MB': os.path.getsize(path) / 1e6, 'users': a.num_users,
                      'items': a.num_items, 'native_s': t_native, 'python_loop_s': t_python, 'identical': bool(same),
                      'cores': cores}))
    os.remove(path)
    n = args.pairs
    U, I = n // 50, 100_000
    u = np.repeat(np.arange(U), 50); rng.shuffle(u)
    i = rng.integers(0, I, len(u))
    t0 = time.perf_counter(); csr = E.RatedCSR(U, I, u, i); t_native = time.perf_counter() - t0
    t0 = time.perf_counter(); ref = O.rated_csr_numpy(U, I, u, i); t_numpy = time.perf_counter() - t0
    same = all(np.array_equal(getattr(csr, k), ref[k]) for k in ref)
    print(json.dumps({'what': 'per-user item sets (rejection CSR + positives in insertion order)', 'pairs': len(u),
                      'users': U, 'items': I, 'native_s': t_native, 'numpy_s': t_numpy, 'identical': bool(same), 'cores': cores}))


if __name__ == '__main__':
    main()
