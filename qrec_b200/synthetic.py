"""Synthetic interaction sets of BASELINE.json's shape, built on the device.

Every user gets exactly `degree` DISTINCT items drawn uniformly (or Zipf-like with
item = floor(I * x^2), x ~ U[0,1)) so the result is a CSR by construction: user-major (u, i)
pairs, plus per-user item rows sorted ascending (the rejection set of the negative sampler).
torch is used here only as a tensor allocator / data generator -- this is test and bench input
preparation, not the hot path.
"""
import torch


def make_interactions(num_users, num_items, degree, seed=20240229, device='cuda', zipf=False,
                      user_offset=0, chunk_users=131072):
    """Returns dict(u=int32[n], i=int32[n] (user-major, random order inside a user),
    sorted_rowptr=int64[num_users+1], sorted_cols=int32[n] (ascending inside a user))."""
    assert degree < num_items
    g = torch.Generator(device=device)
    g.manual_seed(seed + user_offset)
    over = degree + max(8, degree // 4)         # candidates per user; duplicates are discarded
    rows_sorted, rows_rand = [], []
    for start in range(0, num_users, chunk_users):
        m = min(chunk_users, num_users - start)
        if zipf:
            x = torch.rand(m, over, device=device, generator=g)
            cand = (x * x * num_items).to(torch.int32).clamp_(max=num_items - 1)
        else:
            cand = torch.randint(0, num_items, (m, over), device=device, generator=g, dtype=torch.int32)
        s, _ = cand.sort(dim=1)
        dup = torch.zeros_like(s, dtype=torch.bool)
        dup[:, 1:] = s[:, 1:] == s[:, :-1]
        key = torch.rand(m, over, device=device, generator=g)
        key[dup] = 2.0                           # duplicates lose
        sel = key.topk(degree, dim=1, largest=False)
        assert bool((sel.values < 1.5).all()), 'not enough distinct candidates; raise `over`'
        picked = torch.gather(s, 1, sel.indices)                 # random order inside the user
        rows_rand.append(picked)
        rows_sorted.append(picked.sort(dim=1).values)
    items_rand = torch.cat(rows_rand).reshape(-1).contiguous()
    items_sorted = torch.cat(rows_sorted).reshape(-1).contiguous()
    users = torch.arange(num_users, device=device, dtype=torch.int32).repeat_interleave(degree)
    rowptr = torch.arange(num_users + 1, device=device, dtype=torch.int64) * degree
    return dict(u=users, i=items_rand, sorted_rowptr=rowptr, sorted_cols=items_sorted,
                num_users=num_users, num_items=num_items, degree=degree)


def init_tables(num_users, num_items, d, seed=1, device='cuda'):
    """P = U[0,1/3), Q = U[0,1/3) in fp32 -- the distribution of
    base/iterativeRecommender.py:37-38 (np.random.rand(...)/3)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    P = torch.rand(num_users, d, device=device, generator=g) / 3
    Q = torch.rand(num_items, d, device=device, generator=g) / 3
    return P, Q


def build_norm_adj(data, num_users, num_items, device):
    """D^-1/2 (R (+) R^T) D^-1/2 as device CSR (base/graphRecommender.py:10-29), from the
    synthetic user-major pairs.  Setup code (torch ops), not the measured path."""
    u = data['u'].long()
    i_sorted = data['sorted_cols'].long()            # per-user ascending
    n = num_users + num_items
    deg_u = torch.bincount(u, minlength=num_users).double()
    deg_i = torch.bincount(i_sorted, minlength=num_items).double()
    # user rows: already CSR, columns offset by num_users
    vals_u = (1.0 / torch.sqrt(deg_u[u] * deg_i[i_sorted])).float()
    # item rows: sort pairs by (item, user)
    key = i_sorted * num_users + u
    order = torch.argsort(key)
    it, us = i_sorted[order], u[order]
    vals_i = vals_u[order]
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
    rowptr[1:num_users + 1] = torch.cumsum(deg_u.long(), 0)
    rowptr[num_users + 1:] = rowptr[num_users] + torch.cumsum(deg_i.long(), 0)
    cols = torch.cat([(i_sorted + num_users).int(), us.int()]).contiguous()
    vals = torch.cat([vals_u, vals_i]).contiguous()
    return rowptr, cols, vals


