"""Runs the device sampler's SOURCE (qrec_b200/csrc/philox.cuh) on the CPU: the header is compiled with
g++ through tests/host_shims/philox_host.cpp (three intrinsics mapped to plain C++) and compared with the
numpy oracle -- Philox4x32-10 itself, the rejection sampler, and the signature pre-test variant, whose
claim is that it returns exactly the same negatives."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def host(tmp_path_factory):
    out = str(tmp_path_factory.mktemp('shim') / 'libphilox_host.so')
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-I', os.path.join(ROOT, 'qrec_b200', 'csrc'),
                           os.path.join(ROOT, 'tests', 'host_shims', 'philox_host.cpp'), '-o', out])
    lib = C.CDLL(out)
    u32p, i32p, i64p = C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    lib.host_philox4x32_10.argtypes = [C.c_uint32] * 6 + [u32p]
    lib.host_sample_negatives.argtypes = [C.c_int64, C.c_int64, i32p, i64p, i32p, u32p, C.c_int32, C.c_uint64,
                                          C.c_uint32, i32p, i32p]
    return lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def test_philox_known_answers(host):
    # Random123 known-answer vectors for philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    out = (C.c_uint32 * 4)()
    for ctr, key, want in kat:
        host.host_philox4x32_10(*ctr, *key, out)
        assert tuple(out) == want


@pytest.mark.parametrize('num_items,max_deg', [(50, 45), (700, 60), (5000, 600), (100000, 50)])
def test_device_sampler_source_equals_oracle_with_and_without_signature(host, num_items, max_deg):
    from oracle import bpr_oracle as O
    rng = np.random.default_rng(num_items)
    nu = 300
    deg = rng.integers(0, max_deg + 1, nu)
    deg[0] = 0
    rows = [np.sort(rng.choice(num_items, k, replace=False)).astype(np.int32) for k in deg]
    rowptr = np.zeros(nu + 1, np.int64); rowptr[1:] = np.cumsum(deg)
    cols = np.concatenate(rows).astype(np.int32) if rowptr[-1] else np.zeros(0, np.int32)
    sig = np.zeros((nu, 16), np.uint32)
    for uu, r in enumerate(rows):
        for c in r.tolist():
            sig[uu, (c >> 5) & 15] |= np.uint32(1 << (c & 31))
    n = 4000
    users = rng.integers(0, nu, n).astype(np.int32)
    plain, with_sig = np.empty(n, np.int32), np.empty(n, np.int32)
    seed, epoch = 0x1234567890abcdef, 7
    host.host_sample_negatives(n, 0, _p(users, C.c_int32), _p(rowptr, C.c_int64), _p(cols, C.c_int32),
                               _p(sig, C.c_uint32), num_items, seed, epoch, _p(plain, C.c_int32), _p(with_sig, C.c_int32))
    assert np.array_equal(plain, with_sig)
    want = O.sample_neg_philox(users, [set(r.tolist()) for r in rows], num_items, seed, epoch)
    assert np.array_equal(plain, want)
    rated = [set(r.tolist()) for r in rows]
    assert not any(int(j) in rated[int(uu)] for uu, j in zip(users, plain))


def test_signature_never_hides_a_rated_item(host):
    """Adversarial rows: items congruent mod 512, a full row (saturated user) and a row holding every
    item but one."""
    num_items = 2048
    rows = [np.arange(5, num_items, 512, dtype=np.int32), np.arange(num_items, dtype=np.int32),
            np.delete(np.arange(num_items, dtype=np.int32), 1234), np.zeros(0, np.int32)]
    rowptr = np.zeros(len(rows) + 1, np.int64); rowptr[1:] = np.cumsum([len(r) for r in rows])
    cols = np.concatenate(rows)
    sig = np.zeros((len(rows), 16), np.uint32)
    for uu, r in enumerate(rows):
        for c in r.tolist():
            sig[uu, (c >> 5) & 15] |= np.uint32(1 << (c & 31))
    users = np.repeat(np.arange(len(rows), dtype=np.int32), 200)
    n = len(users)
    plain, with_sig = np.empty(n, np.int32), np.empty(n, np.int32)
    host.host_sample_negatives(n, 10**10, _p(users, C.c_int32), _p(rowptr, C.c_int64), _p(cols, C.c_int32),
                               _p(sig, C.c_uint32), num_items, 99, 1, _p(plain, C.c_int32), _p(with_sig, C.c_int32))
    assert np.array_equal(plain, with_sig)
    assert (plain[users == 2] == 1234).all()                 # the only unrated item
    assert not np.isin(plain[users == 0], rows[0]).any()
