#!/usr/bin/env python
"""One-minibatch step times of SimGCL, NGCF and NeuMF on the synthetic 1M x 100K x 50M set (d=64,
reference batch sizes), single GPU.  The model classes are the drop-in ones; only the data object
is replaced by the synthetic generator's arrays.  One JSON line per model."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class FakeData(object):
    def __init__(self, U, I):
        self.user, self.item = range(U), range(I)


def shell(cls, U, I, d, dev, adj, **attrs):
    class Shell(cls):
        def __init__(self):
            pass

        def create_joint_sparse_adj_tensor(self):
            return adj
    m = Shell()
    m.data = FakeData(U, I)
    m.num_users, m.num_items, m.emb_size, m.batch_size = U, I, d, 2048
    m.lRate, m.regU, m.regI, m.engine_device, m.engine_seed = 0.001, 0.001, 0.001, dev.index or 0, 0
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--scale', type=float, default=1.0)
    args = ap.parse_args()
    import torch
    from qrec_b200 import engine as E, synthetic
    from qrec_b200.model.ranking.SimGCL import SimGCL
    from qrec_b200.model.ranking.NGCF import NGCF
    from qrec_b200.model.ranking.NeuMF import NeuMF
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    U, I, DEG, D = int(1_000_000 * args.scale), int(100_000 * args.scale), 50, 64
    data = synthetic.make_interactions(U, I, DEG, device=dev)
    rp, co, va = synthetic.build_norm_adj(data, U, I, dev)

    from qrec_b200.base.graphRecommender import DeviceCSR

    def Adj(split=True):             # the operand class the drop-in models build in initModel (matmul + the row-list products)
        return DeviceCSR.from_tensors((U + I, U + I), rp, co, va, split_row=U if split else None)
    g = torch.Generator(device=dev); g.manual_seed(0)
    idx = torch.randperm(U * DEG, device=dev, generator=g)[:2048]
    bu, bi = data['u'][idx].contiguous(), data['i'][idx].contiguous()
    bj = E.sample_neg_philox(bu, data['sorted_rowptr'], data['sorted_cols'], I, 1, 0)

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / args.steps
    steps_per_epoch = -(-U * DEG // 2048)
    # ---- LightGCN drop-in class (3 layers): one launch over the joint operator vs one launch per bipartite half
    from qrec_b200.model.ranking.LightGCN import LightGCN
    for split in (False, True):
        m = shell(LightGCN, U, I, D, dev, Adj(split), n_layers=3)
        m.initModel()
        ms = timed(lambda: m.train_step(bu, bi, bj))
        print(json.dumps({'model': 'LightGCN', 'n_layers': 3, 'batch': 2048, 'spmm_launch_per_half': split, 'step_ms': ms,
                          'epoch_s': ms * steps_per_epoch / 1e3, 'loss': float(m._loss.item())}), flush=True)
        del m
        torch.cuda.empty_cache()
    # ---- SimGCL (n_layer 2: 3 whole-graph + 3 row-list products forward, 1 scatter + 1 whole-graph backward, noise,
    #      InfoNCE on the batch's unique rows)
    m = shell(SimGCL, U, I, D, dev, Adj(), cl_rate=0.5, eps=0.1, n_layers=2)
    m.initModel()
    ms = timed(lambda: m.train_step(bu, bi, bj))
    print(json.dumps({'model': 'SimGCL', 'n_layers': 2, 'batch': 2048, 'step_ms': ms, 'epoch_s': ms * steps_per_epoch / 1e3,
                      'losses': m.losses()}))
    del m
    torch.cuda.empty_cache()
    # ---- NGCF (2 layers: 2 fwd + 2 bwd SpMM, 8 [N,64]x[64,64] products, 4 weight-gradient products)
    m = shell(NGCF, U, I, D, dev, Adj())
    m.initModel()
    ms = timed(lambda: m.train_step(bu, bi, bj))
    print(json.dumps({'model': 'NGCF', 'batch': 2048, 'step_ms': ms, 'epoch_s': ms * steps_per_epoch / 1e3,
                      'loss': float(m._loss.item())}))
    del m
    torch.cuda.empty_cache()
    # ---- NeuMF (B = 5 * 2048 samples; phase 2 = fused head: everything trains, dense Adam on 4 tables)
    m = shell(NeuMF, U, I, D, dev, None)
    m.initModel()
    pu = bu.repeat_interleave(5).contiguous()
    pi = torch.randint(0, I, (pu.shape[0],), device=dev, generator=g, dtype=torch.int32)
    pi[::5] = bi
    pr = torch.zeros(pu.shape[0], device=dev); pr[::5] = 1.0
    for mode in (0, 1, 2):
        ms = timed(lambda: m.train_step(mode, pu, pi, pr))
        print(json.dumps({'model': 'NeuMF', 'phase': mode, 'samples_per_step': int(pu.shape[0]), 'step_ms': ms,
                          'samples_per_s': pu.shape[0] / ms * 1e3, 'loss': float(m._loss.item())}))


if __name__ == '__main__':
    main()
