#!/usr/bin/env python
"""Small launch programs for ncu captures (one target per run keeps the replay time short):

  python tools/ncu_targets.py k1        fused user-major epoch, BASELINE config 2 (1M x 100K x 50M, d=64)
  python tools/ncu_targets.py k1_hbm    the same kernel on a 1M-item table (256 MB > L2): the HBM-bound regime
  python tools/ncu_targets.py k1_sig    fused epoch with the signature pre-test in the sampler
  python tools/ncu_targets.py rowops    the row gather / scatter-add microbenchmark (L2-resident table)
  python tools/ncu_targets.py spmm      one whole-graph SpMM (1.1M rows, 100M nnz, d=64)
  python tools/ncu_targets.py topn      K8 on 65536 users x 100K items, N=10
  python tools/ncu_targets.py neumf | lightgcn   the bench sections (for launch lists)
Numbers printed under ncu are never bench values."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from qrec_b200 import engine as E, synthetic
    what = sys.argv[1] if len(sys.argv) > 1 else 'k1'
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    U, I, DEG, D = 1_000_000, 100_000, 50, 64
    loss = torch.zeros(1, dtype=torch.float64, device=dev)
    if what in ('k1', 'k1_hbm', 'k1_sig'):
        items = 1_000_000 if what == 'k1_hbm' else I
        data = synthetic.make_interactions(U, items, DEG, device=dev)
        P, Q = synthetic.init_tables(U, items, D, seed=1, device=dev)
        sig = E.rated_signature(data['sorted_rowptr'], data['sorted_cols']) if what == 'k1_sig' else None
        for ep in range(reps):
            if sig is None:
                E.bpr_epoch_usermajor(P, Q, data['sorted_rowptr'], data['i'], data['sorted_rowptr'], data['sorted_cols'], items, 2024, ep,
                                      0.01, 0.001, 0.001, loss)
            else:
                E.bpr_epoch_usermajor_sig(P, Q, data['sorted_rowptr'], data['i'], data['sorted_rowptr'], data['sorted_cols'], sig, items,
                                          2024, ep, 0.01, 0.001, 0.001, loss)
    elif what == 'rowops':
        T = torch.rand(I, 64, device=dev)
        for mode in (0, 1, 2):
            for _ in range(reps):
                E.ubench_row_ops(T, 100_000_000, mode)
    elif what == 'spmm':
        data = synthetic.make_interactions(U, I, DEG, device=dev)
        rp, co, va = synthetic.build_norm_adj(data, U, I, dev)
        X = torch.randn(U + I, D, device=dev) * 0.01
        Y = torch.empty_like(X)
        for _ in range(reps):
            E.spmm_csr(rp, co, va, X, Y, rowsplit=True)
    elif what == 'topn':
        data = synthetic.make_interactions(65536, I, DEG, device=dev)
        P, Q = synthetic.init_tables(65536, I, D, seed=1, device=dev)
        users = torch.arange(65536, dtype=torch.int32, device=dev)
        for _ in range(reps):
            E.score_topn(P, Q, users, data['sorted_rowptr'], data['sorted_cols'], 10)
    elif what in ('neumf', 'lightgcn'):
        import bench
        data = synthetic.make_interactions(U, I, DEG, device=dev)
        if what == 'neumf':
            print(bench.neumf_section(torch, E, data, dev, 6540.5, steps=2, warmup=1))
        else:
            import torch.distributed as dist
            print(bench.lightgcn_section(torch, dist, E, synthetic, data, dev, 6540.5, 0, 1, steps=2, warmup=1))
    else:
        raise SystemExit('unknown target ' + what)
    torch.cuda.synchronize()
    print('done', what, float(loss.item()))


if __name__ == '__main__':
    main()
