#!/usr/bin/env python
"""Row-sharded item table BPR (K7) under torchrun: parity against the single-GPU fused kernel on a
conflict-free batch, then throughput on the synthetic 1M x 100K set with Q sharded over the ranks.

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dist_bpr_sharded.py
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--batch', type=int, default=1 << 21, help='triples per rank per epoch() call')
    ap.add_argument('--minibatch', type=int, default=1 << 18, help='triples per minibatch (one id/row/delta exchange each)')
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from qrec_b200 import engine as E, synthetic, parallel
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('WORLD_SIZE', 1), ('LOCAL_RANK', 0)))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev, pg_options=dist.ProcessGroupNCCL.Options(is_high_priority_stream=True))
    U, I, D, DEG = 1_000_000, 100_000, 64, 50
    lo, hi = parallel.user_range(rank, world, U)
    bi = I // world
    # ---- parity: global conflict-free batch, every rank also runs the single-GPU kernel on full tables
    g = torch.Generator(device=dev); g.manual_seed(5)
    Pf, Qf = synthetic.init_tables(U, I, D, seed=9, device=dev)
    n = 40000
    u = torch.randperm(U, device=dev, generator=g)[:n].int()
    items = torch.randperm(I, device=dev, generator=g)[:2 * n].int()
    i, j = items[:n].contiguous(), items[n:].contiguous()
    Pl, Ql = Pf[lo:hi].clone(), Qf[rank * bi:(rank + 1) * bi].clone()
    loss = torch.zeros(1, dtype=torch.float64, device=dev)
    E.bpr_sgd_batch(Pf, Qf, u, i, j, 0.05, 0.01, 0.01, loss)
    lu, li, lj = parallel.shard_triples_by_user(u, i, j, rank, world, U)
    m = parallel.ShardedItemTableBPR(Pl, Ql, I, rank, world, 0.05, 0.01, 0.01, max_batch=max(args.minibatch, 1 << 16))
    m.epoch(lu, li, lj, batch=7000)           # several minibatches through the two lanes
    torch.cuda.synchronize()
    m.check()
    torch.testing.assert_close(Pl, Pf[lo:hi], rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(Ql, Qf[rank * bi:(rank + 1) * bi], rtol=1e-6, atol=1e-7)
    if rank == 0:
        print(json.dumps({'parity': 'row-sharded Q (all-to-all) == single-GPU fused kernel', 'world': world}))
    del Pf, Qf
    # ---- throughput: each rank trains `batch` of its own users' triples per step
    data = synthetic.make_interactions(hi - lo, I, DEG, device=dev, user_offset=lo)
    perm = torch.randperm((hi - lo) * DEG, device=dev, generator=g)[:args.batch]
    bu, bi_ = data['u'][perm].contiguous(), data['i'][perm].contiguous()
    bj = E.sample_neg_philox(bu, data['sorted_rowptr'], data['sorted_cols'], I, 3, 0)
    m.lr, m.reg_u, m.reg_i = 0.01, 0.001, 0.001
    # user-major order inside the sample, like an epoch of the rank's shard
    order = torch.argsort(bu.long(), stable=True)
    bu, bi_, bj = bu[order].contiguous(), bi_[order].contiguous(), bj[order].contiguous()
    for _ in range(3):
        m.epoch(bu, bi_, bj, args.minibatch)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        m.epoch(bu, bi_, bj, args.minibatch)
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    m.check()
    if rank == 0:
        ms = float(t.item())
        print(json.dumps({'sharded_q_step_ms': ms, 'minibatch': args.minibatch, 'bucket_capacity': m.capacity(args.minibatch), 'world': world, 'triples_per_rank_per_step': args.batch,
                          'triples_per_s_total': args.batch * world / ms * 1e3,
                          'nvlink_bytes_per_triple_each_way': 2 * 4 * D * (world - 1) / world}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
