#!/usr/bin/env python
"""Data-parallel NeuMF under torchrun: user tables row-sharded, item tables + MLP replicated
(parallel.make_user_sharded_neumf).  (1) parity with the single-GPU drop-in class over the three training phases on a
down-scaled problem (same parameters after every step); (2) step time on BASELINE config 4 (1M x 100K, d=64).

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dist_neumf.py
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build(cls, U, I, D, dev, seed=0, widths=None):
    import bench

    class FakeData(object):
        user, item = range(U), range(I)
    m = cls.__new__(cls)
    m.data = FakeData()
    m.num_users, m.num_items, m.emb_size, m.batch_size = U, I, D, 2048
    m.lRate, m.regU, m.regI, m.engine_device, m.engine_seed, m.device = 0.001, 0.001, 0.001, dev.index or 0, seed, dev
    if widths:
        m.mlp_widths = widths
    bench._neumf_init(m)
    return m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--batch', type=int, default=2048)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from qrec_b200 import parallel
    from qrec_b200.model.ranking.NeuMF import NeuMF
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('WORLD_SIZE', 1), ('LOCAL_RANK', 0)))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    Sharded = parallel.make_user_sharded_neumf(NeuMF)
    D = 64
    g = torch.Generator(device=dev); g.manual_seed(5)

    def batch(U, I, B):
        u = torch.randint(0, U, (B // 5,), device=dev, generator=g, dtype=torch.int32).repeat_interleave(5).contiguous()
        i = torch.randint(0, I, (B,), device=dev, generator=g, dtype=torch.int32)
        r = torch.zeros(B, device=dev); r[::5] = 1.0
        return u, i, r

    # ---------------- parity on a small problem: every rank also runs the single-GPU class
    U, I = 4000 * world, 3000
    ref = build(NeuMF, U, I, D, dev)
    for k in ('PG', 'QG', 'PM', 'QM'):
        ref.params[k].mul_(8.0)                          # make the MLP do something visible (as tests/test_gpu_models.py)
    lo, hi = parallel.user_range(rank, world, U)
    m = build(Sharded, hi - lo, I, D, dev).shard(lo)
    for k, v in ref.params.items():
        m.params[k].copy_(v[lo:hi] if k in ('PG', 'PM') else v)
    def resync():
        """The sharded replica takes the reference's parameters and Adam slots (its shard of the user tables): every
        step is then compared at IDENTICAL parameters -- Adam's first steps turn a gradient that rounds to the other
        side of zero into a step of the other sign, and the next step's TF32 products / ReLU masks would inherit it."""
        cut = lambda k, v: v[lo:hi] if k in ('PG', 'PM') else v         # noqa: E731
        for k, v in ref.params.items():
            m.params[k].copy_(cut(k, v))
        for md in (0, 1, 2):
            m.opt_step[md] = ref.opt_step[md]
            for k, (mm, vv) in ref.opt_state[md].items():
                m.opt_state[md][k][0].copy_(cut(k, mm))
                m.opt_state[md][k][1].copy_(cut(k, vv))

    for mode in (0, 1, 2):
        for step in range(3):
            u, i, r = batch(U, I, 2560)
            resync()
            l_ref = float(ref.train_step(mode, u, i, r).item())
            l = float(m.train_step(mode, u, i, r).item())
            assert abs(l - l_ref) <= 1e-4 * abs(l_ref), (mode, step, l, l_ref)
            # gradients (after the reduction): the sums over samples are regrouped by rank, so fp32 / TF32 rounding
            # differs -- compare against the largest entry; then the parameters: Adam turns a gradient that rounds
            # to the other side of zero into a step of the other sign, so a parameter may differ by up to 2 lr
            for k in m.opt_vars[mode]:
                gr = ref.grads[k][lo:hi] if k in ('PG', 'PM') else ref.grads[k]
                assert float((m.grads[k] - gr).abs().max()) <= 2e-3 * float(ref.grads[k].abs().max()) + 1e-7, (k, mode, step)
            for k, v in ref.params.items():
                mine = v[lo:hi] if k in ('PG', 'PM') else v
                # (one step from identical parameters: at most one sign flip per entry)
                assert float((m.params[k] - mine).abs().max()) <= 2.1 * ref.lRate, (k, mode, step, 'max')
                assert float((m.params[k] - mine).abs().mean()) <= 0.1 * ref.lRate, (k, mode, step, 'mean')
    if world > 1:                                          # replicated parameters stay bit-identical across ranks
        for k in ('QG', 'W1', 'h_mlp'):
            parts = [torch.empty_like(m.params[k]) for _ in range(world)]
            dist.all_gather(parts, m.params[k])
            assert all(torch.equal(parts[0], t) for t in parts), k
    if rank == 0:
        print(json.dumps({'parity': 'UserShardedNeuMF == single-GPU NeuMF over 3 phases x 3 steps, each from identical parameters and Adam slots (losses 1e-4, '
                                    'reduced gradients 2e-3 of max, parameters within one Adam sign flip; replicas bit-identical)', 'world': world, 'problem': [U, I]}))
    del ref, m
    torch.cuda.empty_cache()

    # ---------------- timing at config 4
    U, I = 1_000_000, 100_000
    lo, hi = parallel.user_range(rank, world, U)
    out = {'world': world, 'samples_per_step': 5 * args.batch, 'workload': 'NeuMF 1M x 100K, d=64, reference MLP widths'}
    for name, widths in (('reference_2d_5d_2d_d', None), ('baseline_256_128_64', (256, 128, 64))):
        m = build(Sharded, hi - lo, I, D, dev, widths=widths).shard(lo)
        sec = {}
        for mode, label in ((0, 'gmf'), (1, 'mlp'), (2, 'neumf')):
            batches = [batch(U, I, 5 * args.batch) for _ in range(args.steps + 3)]
            for t in range(3):
                m.train_step(mode, *batches[t])
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for t in range(args.steps):
                m.train_step(mode, *batches[3 + t])
            b.record()
            torch.cuda.synchronize()
            tt = torch.tensor([a.elapsed_time(b) / args.steps], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            sec[label] = {'ms_per_step': float(tt.item()), 'samples_per_s': 5 * args.batch / float(tt.item()) * 1e3}
        out[name] = sec
        del m
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
