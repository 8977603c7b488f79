"""world_size-2 data-parallel test on CPU (gloo): sharding + gradient scaling + ONE all-reduce of the packed flat
buffer (gradients | cost tail) reproduces the single-process full-batch gradient of mean(cost) (nats.py:1323, 1340).
The per-rank gradients come from the oracle (the product needs a GPU); what is under test is the host-side
data-parallel logic of nats_b200.parallel and the packed layout of nats_b200.layout."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nats_b200 import _lib, layout, parallel
from oracle import nats_oracle as O
from tests.helpers import toy_options, toy_params


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    opts = toy_options(D=8, W=6, A=5, V=50)
    P = toy_params(opts)
    rng = np.random.RandomState(3)
    sx = [list(rng.randint(2, 50, size=rng.randint(2, 9))) for _ in range(6)]
    sy = [list(rng.randint(2, 50, size=rng.randint(1, 6))) for _ in range(6)]
    return opts, P, sx, sy


def _worker(rank, world_size, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world_size)
    try:
        opts, P, sx, sy = _problem()
        assert parallel.world() == (rank, world_size)
        lx, ly = parallel.shard(sx, sy, rank, world_size)
        x, xm, y, ym = O.prepare_data(lx, ly, n_words=50)          # each shard pads to ITS longest sentence
        cost_b, cache = O.model_fwd(P, x, xm, y, ym)
        scale = parallel.grad_scale(len(lx), world_size, global_batch=len(sx))
        G = O.model_bwd(P, cache, np.full((len(lx),), scale))
        flat = layout.pack(G, tail=_lib.GRAD_TAIL)
        total = flat.size - _lib.GRAD_TAIL
        flat[total] = scale * cost_b.sum()                          # cost travels in the tail slot
        t = torch.from_numpy(flat)
        parallel.allreduce_flat(t)
        np.save(os.path.join(out_dir, 'rank%d.npy' % rank), t.numpy())
    finally:
        dist.destroy_process_group()


def test_dp2_allreduce_equals_full_batch(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / 'rank0.npy')
    r1 = np.load(tmp_path / 'rank1.npy')
    np.testing.assert_array_equal(r0, r1)                           # replicas see identical buffers
    opts, P, sx, sy = _problem()
    x, xm, y, ym = O.prepare_data(sx, sy, n_words=50)
    mean_cost, G, _ = O.f_grad(P, x, xm, y, ym)
    total = r0.size - _lib.GRAD_TAIL
    assert abs(r0[total] - mean_cost) < 1e-5 * abs(mean_cost)
    got = layout.unpack(r0[:total], layout.dims_of(P))
    for k in G:
        np.testing.assert_allclose(got[k], G[k], rtol=2e-5, atol=1e-7, err_msg=k)
    # padding of the packed layout stays zero (so the global norm over the flat buffer equals the reference's g2)
    mask = np.ones(total, bool)
    views, _ = _lib.param_layout(*layout.dims_of(P))
    for (name, off, rows, cols, ld, ndim) in views:
        idx = (off + np.arange(rows)[:, None] * ld + np.arange(cols)[None, :]).ravel()
        mask[idx] = False
    assert np.all(r0[:total][mask] == 0)
    g2_flat = float((r0[:total].astype('float64') ** 2).sum())
    g2_ref = float(sum((g ** 2).sum() for g in G.values()))
    assert abs(g2_flat - g2_ref) < 1e-5 * g2_ref


def test_shard_and_scale():
    sx = [[1]] * 7
    sy = [[2]] * 7
    parts = [parallel.shard(sx, sy, r, 3) for r in range(3)]
    assert [len(p[0]) for p in parts] == [3, 3, 1]
    assert parallel.grad_scale(32, 8) == 1.0 / 256
    t = torch.ones(4)
    assert parallel.allreduce_flat(t) is t                           # single process: no-op


def test_pack_unpack_roundtrip():
    opts = toy_options(D=8, W=6, A=5, V=50)
    P = O.cast_params(toy_params(opts), 'float32')
    flat = layout.pack(P)
    back = layout.unpack(flat, layout.dims_of(P))
    assert list(back.keys()) == list(P.keys())
    for k in P:
        np.testing.assert_array_equal(back[k], P[k])
