"""Batched top-N evaluation on the device ("next" row).

The reference ranks one test user at a time: an I x d GEMV, rated items overwritten with 0, a
numba heap top-N (base/recommender.py:143-152).  Here a block of users is scored at once:
    scores = U[block] @ V^T        qrec_sgemm_f32 (fp32 SIMT: same ranking as the GEMV to rounding)
    scores[rated] = 0              qrec_mask_rated_f32
    top-N per row                  torch.topk  (library selection kernel; N <= 100)
Opt-in through `engine=... -eval gpu`; the default keeps the reference's host flow, whose float64
score strings are part of the recorded outputs.
"""
import numpy as np


def batched_top_n(U, V, user_ids, csr, N, block=2048):
    """U [users,d], V [items,d]: fp32 CUDA tensors; user_ids: int array of rows of U to rank;
    csr: engine.RatedCSR of the training set.  Returns (ids [n,N] int64, scores [n,N] float32)."""
    import torch
    from . import engine as E
    dev = U.device
    rowptr = torch.from_numpy(csr.sorted_rowptr).to(dev)
    cols = torch.from_numpy(csr.sorted_cols).to(dev)
    user_ids = np.asarray(user_ids, dtype=np.int32)
    n, I = len(user_ids), V.shape[0]
    N = min(N, I)
    out_ids = np.empty((n, N), np.int64)
    out_val = np.empty((n, N), np.float32)
    scores = torch.empty(min(block, max(n, 1)), I, device=dev)
    for b in range(0, n, block):
        ub = torch.from_numpy(user_ids[b:b + block]).to(dev)
        m = ub.shape[0]
        Ub = U.index_select(0, ub.long()).contiguous()
        E.sgemm(Ub, V, scores[:m], trans_b=True)
        E.mask_rated(scores[:m], ub, rowptr, cols, 0.0)
        val, idx = torch.topk(scores[:m], N, dim=1, largest=True, sorted=True)
        out_ids[b:b + m] = idx.cpu().numpy()
        out_val[b:b + m] = val.cpu().numpy()
    return out_ids, out_val
