#!/usr/bin/env python
"""Parity-mode (dependency-ordered) kernel at a tenth of the benchmark scale: 100K users x 10K items
x 5M triples in the reference's user-major order, d=64, fp32 and fp64.  One JSON line each."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from qrec_b200 import engine as E, synthetic
    dev = torch.device('cuda', 0)
    U, I, DEG, D = 100_000, 10_000, 50, 64
    data = synthetic.make_interactions(U, I, DEG, device=dev)
    u, i = data['u'], data['i']
    j = E.sample_neg_philox(u, data['sorted_rowptr'], data['sorted_cols'], I, 1, 0)
    hu, hi, hj = u.cpu().numpy(), i.cpu().numpy(), j.cpu().numpy()
    t0 = time.perf_counter()
    wu, wi, wj = E.bpr_order_prepare(hu, hi, hj, U, I)
    prep = time.perf_counter() - t0
    dw = [torch.from_numpy(x).to(dev) for x in (wu, wi, wj)]
    depth = E.bpr_order_depth(hu, hi, hj, U, I)
    width = len(hu) / depth
    for dt, nw in ((torch.float32, 0), (torch.float32, int(max(32, 4 * width))), (torch.float32, int(max(32, 16 * width))),
                   (torch.float64, int(max(32, 4 * width)))):
        P, Q = synthetic.init_tables(U, I, D, device=dev)
        P, Q = P.to(dt), Q.to(dt)
        loss = torch.zeros(1, dtype=torch.float64, device=dev)
        E.bpr_sgd_ordered(P, Q, u, i, j, *dw, 0.01, 0.001, 0.001, loss, n_warps=nw)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            E.bpr_sgd_ordered(P, Q, u, i, j, *dw, 0.01, 0.001, 0.001, loss, n_warps=nw)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        print(json.dumps({'kernel': 'bpr_sgd_ordered', 'dtype': str(dt), 'triples': int(u.numel()), 'ms': ms,
                          'M_triples_s': u.numel() / ms / 1e3, 'host_prepare_s': prep, 'dag_depth': depth,
                          'dag_width': width, 'n_warps': nw}))


if __name__ == '__main__':
    main()
