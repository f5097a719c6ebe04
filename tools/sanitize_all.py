#!/usr/bin/env python
"""Launches every kernel of libqrec.so once at small sizes -- meant to run under
  compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_all.py
(the reference has no sanitizer story; K1 throughput mode is racy by design
through atomics only, everything else must be clean)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from qrec_b200 import engine as E, parallel
    torch.cuda.set_device(0)
    rng = np.random.default_rng(0)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()       # noqa: E731
    nu, ni, d, n = 300, 400, 64, 1000
    u = rng.integers(0, nu, n).astype(np.int32); i = rng.integers(0, ni, n).astype(np.int32)
    j = ((i + 1 + rng.integers(0, ni - 1, n)) % ni).astype(np.int32)
    P, Q = dev((rng.random((nu, d)) / 3).astype(np.float32)), dev((rng.random((ni, d)) / 3).astype(np.float32))
    loss = torch.zeros(1, dtype=torch.float64, device='cuda')
    csr = E.RatedCSR(nu, ni, u, i)
    jj = E.sample_neg_philox(dev(u), dev(csr.sorted_rowptr), dev(csr.sorted_cols), ni, 1, 0)
    E.bpr_sgd_batch(P, Q, dev(u), dev(i), jj, 0.01, 0.001, 0.001, loss)
    for dd in (8, 48, 128, 256):
        Pd, Qd = torch.rand(nu, dd, device='cuda'), torch.rand(ni, dd, device='cuda')
        E.bpr_sgd_batch(Pd, Qd, dev(u), dev(i), dev(j), 0.01, 0.001, 0.001, loss)
    wu, wi, wj = E.bpr_order_prepare(u, i, j, nu, ni)
    for dt in (torch.float32, torch.float64):
        E.bpr_sgd_ordered(P.to(dt), Q.to(dt), dev(u), dev(i), dev(j), dev(wu), dev(wi), dev(wj), 0.01, 0.001, 0.001, loss)
    E.sumsq(P, loss); E.sumsq(P.double(), loss)
    pipe = E.HostPipeline(0, chunk_triples=300)
    pipe.bpr_epoch(P, Q, u, i, j, 0.01, 0.001, 0.001); pipe.close()
    m = parallel.ShardedItemTableBPR(P, Q, ni, 0, 1, 0.01, 0.001, 0.001)
    m.step(dev(u), dev(i), dev(j))
    # graph path
    import scipy.sparse as sp
    N = nu + ni
    A = sp.random(N, N, density=0.02, format='csr', dtype=np.float32, random_state=1); A.sort_indices()
    rp, co, va = dev(A.indptr.astype(np.int64)), dev(A.indices.astype(np.int32)), dev(A.data)
    X, Y, acc = torch.rand(N, d, device='cuda'), torch.empty(N, d, device='cuda'), torch.zeros(N, d, device='cuda')
    for rs in (False, True):
        E.spmm_csr(rp, co, va, X, Y, acc=acc, acc_scale=0.5, rowsplit=rs)
    gU, gV = torch.zeros_like(P), torch.zeros_like(Q)
    E.bpr_grad_scatter(P, Q, dev(u), dev(i), dev(j), 1e-7, 0.001, gU, gV, loss)
    E.adam_dense_tf1(P, torch.zeros_like(P), torch.zeros_like(P), gU, 0.001, 1)
    E.axpby(Y, X, acc, 1.0, 2.0)
    # K6 / dense
    E.simgcl_perturb(X, 0.1, 7, 1, 1, acc=acc, acc_scale=0.5, d_valid=62)
    idx = dev(rng.permutation(N)[:129].astype(np.int32))
    Z, nrm = torch.empty(129, d, device='cuda'), torch.empty(129, device='cuda')
    E.gather_normalize(X, idx, Z, nrm)
    S = torch.empty(129, 129, device='cuda')
    E.sgemm(Z, Z, S, trans_b=True)
    E.infonce_rows(S, 0.2, loss)
    dZ = torch.empty_like(Z)
    E.sgemm(S, Z, dZ); E.sgemm(S, Z, dZ, trans_a=True)
    E.normalize_bwd_scatter(dZ, Z, nrm, idx, 0.5, acc)
    W = torch.rand(d, d, device='cuda')
    big = torch.rand(5000, d, device='cuda')
    E.sgemm(big, big, W, trans_a=True)                    # split-K path
    H, out, norms = torch.empty(N, d, device='cuda'), torch.empty(N, 3 * d, device='cuda'), torch.empty(N, device='cuda')
    E.ngcf_act_fwd(X, 0.9, 1, 3, 0, 1, H, out[:, d:2 * d], norms)
    E.ngcf_act_bwd(out[:, d:2 * d], None, H, X, norms, 0.9, 1, 3, 0, 1, Y)
    E.mul(Y, X, acc)
    # K5
    B = 300
    A0 = torch.rand(B, 128, device='cuda'); W1 = torch.rand(128, 320, device='cuda'); b1 = torch.rand(320, device='cuda')
    H1 = torch.empty(B, 320, device='cuda')
    E.tc_gemm(A0, W1, H1, epilogue=E.EPI_BIAS_RELU, bias=b1)
    dX = torch.empty(B, 128, device='cuda')
    E.tc_gemm(H1, W1, dX, b_is_nk=True, epilogue=E.EPI_RELU_MASK, mask=A0)
    uu, ii = dev(u[:B]), dev(i[:B])
    X0 = torch.empty(B, 2 * d, device='cuda')
    E.gather_rows(P, uu, X0[:, :d]); E.gather_rows(Q, ii, X0[:, d:])
    E.scatter_add_rows(gU, uu, X0[:, :d])
    y, dz = torch.empty(B, device='cuda'), torch.empty(B, device='cuda')
    UG, IG, H3 = (torch.rand(B, d, device='cuda') for _ in range(3))
    GMF, dUG, dIG, dH3 = (torch.empty(B, d, device='cuda') for _ in range(4))
    hm, hl = torch.rand(d, device='cuda'), torch.rand(d, device='cuda')
    r = (torch.rand(B, device='cuda') > 0.8).float()
    for mode in (0, 1, 2):
        E.neumf_head(mode, 1, UG, IG, H3, hm, hl, r, 0.001, loss, y, dz, GMF, dUG, dIG, dH3)
    sc = torch.rand(64, ni, device='cuda')
    E.mask_rated(sc, dev(np.arange(64, dtype=np.int32)), dev(csr.sorted_rowptr), dev(csr.sorted_cols))
    torch.cuda.synchronize()
    print('sanitize_all: launched', E.launch_count(), 'kernels')
    if True:
        # K9 (rating-prediction MF): kept apart until its first hardware run has passed
        n9 = n
        u9, i9 = np.ascontiguousarray(u[:n9]), np.ascontiguousarray(i[:n9])
        r9 = torch.rand(n9, device='cuda') * 4
        wu9, wi9 = E.mf_order_prepare(u9, i9, nu, ni)
        Bu, Bi = torch.zeros(nu, device='cuda'), torch.zeros(ni, device='cuda')
        for kind in (0, 1, 2):
            E.mf_sgd_batch(kind, P, Q, dev(u9), dev(i9), r9, 0.01, 0.01, 0.01, loss, Bu, Bi, 0.01, 2.0)
            E.mf_sgd_ordered(kind, P, Q, dev(u9), dev(i9), r9, dev(wu9), dev(wi9), 0.01, 0.01, 0.01, loss, Bu, Bi, 0.01, 2.0)
        E.mf_predict_pairs(P, Q, dev(u9), dev(i9), Bu, Bi, 2.0)
        Hv2 = torch.empty(B, 320, device='cuda')
        E.tc_gemm_v2(A0, W1, Hv2, epilogue=E.EPI_BIAS_RELU, bias=b1)
        E.tc_gemm_v2(H1, W1, dX, b_is_nk=True, epilogue=E.EPI_RELU_MASK, mask=A0)
        sig = E.rated_signature(dev(csr.sorted_rowptr), dev(csr.sorted_cols))
        E.bpr_epoch_usermajor_sig(P, Q, dev(csr.pos_rowptr), dev(csr.pos_cols), dev(csr.sorted_rowptr),
                                  dev(csr.sorted_cols), sig, ni, 5, 0, 0.01, 0.001, 0.001, loss)
        torch.cuda.synchronize()
        print('sanitize_all: + K9, launched', E.launch_count(), 'kernels')
        # round-2 kernels
        from qrec_b200.graph_build import JointAdjacency
        E.bpr_epoch_usermajor(P, Q, dev(csr.pos_rowptr), dev(csr.pos_cols), dev(csr.sorted_rowptr), dev(csr.sorted_cols), ni, 5, 0,
                              0.01, 0.001, 0.001, loss)
        E.bpr_epoch_usermajor_tma(P, Q, dev(csr.pos_rowptr), dev(csr.pos_cols), dev(csr.sorted_rowptr), dev(csr.sorted_cols), ni, 5, 0,
                                  0.01, 0.001, 0.001, loss)
        ids, vals = E.score_topn(P, Q, dev(np.arange(nu, dtype=np.int32)), dev(csr.sorted_rowptr), dev(csr.sorted_cols), 10)
        J = JointAdjacency(dev(u.astype(np.int64)), dev(i.astype(np.int64)), nu, ni, device='cuda')
        J.full(); J.edge_dropout(0.3, 1, 2, 3)
        Bt, Dt, St = Q.clone(), torch.empty(Q.numel(), device='cuda'), torch.empty(Q.numel(), device='cuda')
        E.table_delta(Q.view(-1), Bt.view(-1), Dt, St)
        E.table_reduce_scatter_p2p([Dt.data_ptr(), Dt.data_ptr()], 1, St, Dt.numel())
        E.table_gather_merge_p2p([St.data_ptr(), St.data_ptr()], Q.view(-1), Bt.view(-1), Dt)
        E.table_all_gather_p2p([St.data_ptr(), St.data_ptr()], Dt)
        E.table_merge(Q.view(-1), Bt.view(-1), Dt, St)
        E.ubench_row_ops(torch.rand(1000, 64, device='cuda'), 5000, 2)
        cnt, snd = torch.empty(2, dtype=torch.int32, device='cuda'), torch.empty(2 * 900, dtype=torch.int32, device='cuda')
        pos, ovf = torch.empty(n, dtype=torch.int32, device='cuda'), torch.zeros(1, dtype=torch.int32, device='cuda')
        E.bucket_requests(dev(i), ni // 2, 2, 900, cnt, snd, pos, ovf)
        E.simgcl_perturb(X, 0.1, 7, 1, 1, acc=acc, acc_scale=0.5, d_valid=62, row_offset=12345)
        torch.cuda.synchronize()
        print('sanitize_all: + round 2, launched', E.launch_count(), 'kernels')


if __name__ == '__main__':
    main()
