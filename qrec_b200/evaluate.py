"""Batched top-N evaluation on the device ("next" row).

The reference ranks one test user at a time: an I x d GEMV, rated items overwritten with 0, a
numba heap top-N (base/recommender.py:143-152).  Here blocks of users go through ONE kernel
(qrec_score_topn_f32, csrc/topn_kernels.cu): score tile -> compare with the row's N-th best in registers ->
rated test for the survivors -> per-row candidate buffer in shared memory; the [users x items] score
matrix is never written.  Ties are broken by ascending item id (the heap's strict `>` keeps the earlier
item at the cut).  Opt-in through `engine=... -eval gpu`; the default keeps the reference's host flow,
whose float64 score strings are part of the recorded outputs.
"""
import numpy as np


def batched_top_n(U, V, user_ids, csr, N, block=65536):
    """U [users,d], V [items,d]: fp32 CUDA tensors; user_ids: int array of rows of U to rank;
    csr: engine.RatedCSR of the training set.  Returns (ids [n,N] int64, scores [n,N] float32)."""
    import torch
    from . import engine as E
    dev = U.device
    rowptr = torch.from_numpy(csr.sorted_rowptr).to(dev)
    cols = torch.from_numpy(csr.sorted_cols).to(dev)
    user_ids = np.ascontiguousarray(user_ids, dtype=np.int32)
    n, I = len(user_ids), V.shape[0]
    N = min(N, I)
    out_ids = np.empty((n, N), np.int64)
    out_val = np.empty((n, N), np.float32)
    U, V = U.contiguous(), V.contiguous()
    for b in range(0, n, block):
        ub = torch.from_numpy(user_ids[b:b + block]).to(dev)
        ids, val = E.score_topn(U, V, ub, rowptr, cols, N, rated_value=0.0)
        out_ids[b:b + ub.shape[0]] = ids.cpu().numpy()
        out_val[b:b + ub.shape[0]] = val.cpu().numpy()
    return out_ids, out_val
