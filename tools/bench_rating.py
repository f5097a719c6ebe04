#!/usr/bin/env python
"""K9 (rating-prediction MF step) on the benchmark-scale synthetic set: 1M x 100K x 50M ratings, d=64,
shuffled order.  Times qrec_mf_sgd_batch_f32 per kind (full grid and two bounded in-flight windows) and qrec_mf_predict_pairs_f32, plus the C port on a 2 M-entry sample.  One JSON line each.
Algorithmic bytes per entry: 2 rows x (read + reduce) x 4d + 12 B of (u, i, r) = 1036 B at d = 64."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(torch, fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    import torch
    from qrec_b200 import engine as E, synthetic
    dev = torch.device('cuda', 0)
    U, I, DEG, D = 1_000_000, 100_000, 50, 64
    data = synthetic.make_interactions(U, I, DEG, device=dev)
    n = U * DEG
    g = torch.Generator(device=dev); g.manual_seed(1)
    perm = torch.randperm(n, device=dev, generator=g)
    u, i = data['u'][perm].contiguous(), data['i'][perm].contiguous()
    r = (torch.randint(1, 9, (n,), device=dev, generator=g).float() / 2).contiguous()
    Bu = torch.zeros(U, device=dev); Bi = torch.zeros(I, device=dev)
    loss = torch.zeros(1, dtype=torch.float64, device=dev)
    bytes_per_entry = 2 * 2 * 4 * D + 12
    for kind, name in ((0, 'BasicMF'), (1, 'PMF'), (2, 'SVD')):
        P, Q = synthetic.init_tables(U, I, D, device=dev)
        for window in (0, 1 << 16, 1 << 13):
            def epoch():
                E.mf_sgd_batch(kind, P, Q, u, i, r, 1e-4, 0.01, 0.01, loss, Bu if kind == 2 else None,
                               Bi if kind == 2 else None, 0.01, 2.5, max_inflight=window)
            ms = timed(torch, epoch, reps=5, warm=2)
            print(json.dumps({'k9': name, 'max_inflight': window, 'ms_per_50M': ms, 'G_entries_s': n / ms / 1e6,
                              'algorithmic_TBs': n * bytes_per_entry / ms / 1e9,
                              'finite': bool(torch.isfinite(P).all().item())}))
    P, Q = synthetic.init_tables(U, I, D, device=dev)
    out = torch.empty(n, device=dev)
    ms = timed(torch, lambda: E.mf_predict_pairs(P, Q, u, i, out=out))
    print(json.dumps({'k9': 'predict_pairs', 'ms_per_50M': ms, 'G_pairs_s': n / ms / 1e6}))
    # CPU port (oracle/mf_ref.c, one thread like the reference's loop) on a bounded sample
    from oracle import c_oracle
    m = 2_000_000
    Ph, Qh = P.cpu().numpy(), Q.cpu().numpy()
    uh, ih, rh = u[:m].cpu().numpy(), i[:m].cpu().numpy(), r[:m].cpu().numpy()
    t0 = time.perf_counter()
    c_oracle.mf_sgd_sequential(1, Ph, Qh, uh, ih, rh, 1e-4, 0.01, 0.01)
    dt = time.perf_counter() - t0
    print(json.dumps({'k9': 'cpu_port_pmf_f32', 'sample': m, 'M_entries_s': m / dt / 1e6, 'cores': 1}))


if __name__ == '__main__':
    main()
