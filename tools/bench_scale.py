#!/usr/bin/env python
"""f-3 at the benchmark size: a 1M x 100K x 50M InteractionTable (flat arrays, no Python records) through
ScaleBPR -- host-side preparation times, epoch time, loss curve.  One JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from qrec_b200.data.interactions import InteractionTable
    from qrec_b200.scale import ScaleBPR
    U, I, DEG = 1_000_000, 100_000, 50
    rng = np.random.default_rng(0)
    t0 = time.perf_counter()
    u = np.repeat(np.arange(U, dtype=np.int32), DEG)
    i = rng.integers(0, I, U * DEG, dtype=np.int32)          # ~1 % duplicate pairs, as real logs have
    table = InteractionTable(np.arange(U).astype(str), np.arange(I).astype(str), u, i, np.ones(U * DEG))
    t_table = time.perf_counter() - t0
    np.random.seed(0)
    t0 = time.perf_counter()
    m = ScaleBPR(table, emb_size=64, lr=0.01)
    torch.cuda.synchronize()
    t_prep = time.perf_counter() - t0
    m.run_epoch()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        m.run_epoch()
    torch.cuda.synchronize()
    per_epoch = (time.perf_counter() - t0) / 5
    n = int(m.pos_items.numel())
    print(json.dumps({'interactions': len(table), 'positives_after_dedup': n, 'table_build_s': t_table,
                      'csr_and_upload_s': t_prep, 'epoch_ms_wall': per_epoch * 1e3, 'G_triples_s': n / per_epoch / 1e9,
                      'loss_curve': [round(h[1], 1) for h in m.history], 'lr': [h[3] for h in m.history]}))


if __name__ == '__main__':
    main()
