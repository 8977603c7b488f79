"""Oracle self-consistency: f_init/f_next chain == training scan (masks all ones), SciPy pins for the
beam-search distraction penalties (nats.py:990-995), invariants of SURVEY 8(c)(3), optimiser steps."""
import numpy as np
import scipy.spatial.distance
import scipy.stats

from oracle import nats_oracle as O
from tests.helpers import toy_options, toy_params, ragged_batch, full_batch


def test_entropy_cosine_match_scipy():
    rng = np.random.RandomState(3)
    for _ in range(20):
        p = rng.rand(17).astype('float32') + 1e-3
        q = rng.rand(17).astype('float32') + 1e-3
        assert np.isclose(O.entropy(p, q), scipy.stats.entropy(p, q), rtol=1e-6)
        u = rng.randn(33).astype('float32'); v = rng.randn(33).astype('float32')
        assert np.isclose(O.cosine(u, v), scipy.spatial.distance.cosine(u, v), rtol=1e-5, atol=1e-6)
    p = np.array([0., 0.5, 0.5]); q = np.array([0.2, 0.3, 0.5])
    assert np.isclose(O.entropy(p, q), scipy.stats.entropy(p, q))


def test_sampler_chain_equals_scan():
    """with all-ones masks, f_init + chained f_next reproduces the scan outputs."""
    opts = toy_options()
    P = toy_params(opts)
    x, xm, y, ym = full_batch(opts['n_words'], B=2, Tx=7, Ty=5)
    cost, cache = O.model_fwd(P, x, xm, y, ym)
    init_state, ctx = O.f_init(P, x)
    np.testing.assert_allclose(ctx, cache['ctx'], rtol=1e-13)
    np.testing.assert_allclose(init_state, cache['init_state'], rtol=1e-13)
    B = x.shape[1]
    state = init_state
    acc_ctx = np.zeros((B, ctx.shape[2])); acc_alpha = np.zeros((B, ctx.shape[0]))
    yprev = -np.ones((B,), 'int64')
    nll = np.zeros(B)
    for t in range(y.shape[0]):
        probs, _, state, alT, c, acc_ctx, acc_alpha = O.f_next(P, yprev, ctx, state, acc_ctx, acc_alpha)
        np.testing.assert_allclose(state, cache['Hs'][t], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(c, cache['Cs'][t], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(alT, cache['As'][t], rtol=1e-12, atol=1e-14)
        nll += -np.log(probs[np.arange(B), y[t]])
        yprev = y[t]
    np.testing.assert_allclose(nll, cost, rtol=1e-11)


def test_attention_invariants_and_padding_invariance():
    opts = toy_options()
    P = toy_params(opts)
    x, xm, y, ym = ragged_batch(opts['n_words'])
    cost, cache = O.model_fwd(P, x, xm, y, ym)
    As = cache['As']                                        # [Ty,B,Tx]
    np.testing.assert_allclose(As.sum(2), 1.0, rtol=1e-12)  # rows sum to one (:540)
    assert np.all(As * (1 - xm.T[None]) == 0)               # vanish on padded positions (:538-539)
    _, _, _, acc_ctx, acc_alpha = O.gru_cond_layer_fwd(P, cache['embs'], cache['ym'], cache['ctx'],
                                                        cache['xm'], cache['init_state'])[0]
    np.testing.assert_allclose(acc_alpha.sum(1), ym.sum(0), rtol=1e-12)   # (:570)
    # extra padding rows/cols change neither cost nor gradients
    pad = lambda a, n: np.concatenate([a, np.zeros((n,) + a.shape[1:], a.dtype)], 0)
    cost2 = O.f_log_probs(P, pad(x, 3), pad(xm, 3), pad(y, 2), pad(ym, 2))
    np.testing.assert_allclose(cost2, cost, rtol=1e-12)
    _, G, _ = O.f_grad(P, x, xm, y, ym)
    _, G2, _ = O.f_grad(P, pad(x, 3), pad(xm, 3), pad(y, 2), pad(ym, 2))
    for k in G:
        np.testing.assert_allclose(G2[k], G[k], rtol=1e-9, atol=1e-13)


def test_optimizers_one_step():
    opts = toy_options()
    P = toy_params(opts)
    x, xm, y, ym = ragged_batch(opts['n_words'])
    _, G, _ = O.f_grad(P, x, xm, y, ym)
    # adadelta, first step from zero state: ud = -sqrt(eps)/sqrt(0.05 g^2 + eps) * g   (:1156-1168)
    P1 = O.cast_params(P, 'float64'); opt = O.Adadelta(P1)
    opt.grad_shared(G); opt.update(P1)
    for k in P:
        ud = -np.sqrt(1e-6) / np.sqrt(0.05 * G[k] ** 2 + 1e-6) * G[k]
        np.testing.assert_allclose(P1[k], P[k] + ud, rtol=1e-12, atol=1e-15)
    # adam first step (:1114-1136): m=b1 g, v=b2 g^2, lr_t = lr0 sqrt(1-b2)/(1-b1)
    P2 = O.cast_params(P, 'float64'); opt = O.Adam(P2)
    opt.grad_shared(G); opt.update(P2)
    lr_t = 0.0002 * np.sqrt(1 - 0.001) / (1 - 0.1)
    for k in P:
        np.testing.assert_allclose(P2[k], P[k] - lr_t * (0.1 * G[k]) / (np.sqrt(0.001 * G[k] ** 2) + 1e-8),
                                   rtol=1e-12, atol=1e-15)
    P3 = O.cast_params(P, 'float64'); opt = O.RMSprop(P3)
    opt.grad_shared(G); opt.update(P3)
    for k in P:
        ud = -1e-4 * G[k] / np.sqrt(0.05 * G[k] ** 2 - (0.05 * G[k]) ** 2 + 1e-4)
        np.testing.assert_allclose(P3[k], P[k] + ud, rtol=1e-12, atol=1e-15)


def _callables(P):
    fi = lambda x: O.f_init(P, x)
    fn = lambda y, ctx, s, ac, aa: O.f_next(P, y, ctx, s.astype(P['Wemb'].dtype), ac.astype(P['Wemb'].dtype),
                                            aa.astype(P['Wemb'].dtype))
    return fi, fn


def test_beam_search_runs_and_penalties_change_ranking():
    opts = toy_options(V=30)
    P = toy_params(opts, dtype='float32')
    fi, fn = _callables(P)
    x = np.array([3, 7, 9, 4, 11, 5, 0], 'int64')[:, None]
    s0, sc0, al0 = O.gen_sample(fi, fn, x, k=3, maxlen=8, stochastic=False, use_unk=True)
    assert len(s0) == len(sc0) == len(al0) and 1 <= len(s0) <= 3
    tr = []
    s1, sc1, _ = O.gen_sample(fi, fn, x, k=3, maxlen=8, stochastic=False, use_unk=True,
                              kl_factor=2.0, ctx_factor=2.0, state_factor=2.0, trace=tr)
    assert any(t['pen'] is not None for t in tr)
    pen = [t['pen'] for t in tr if t['pen'] is not None][0]
    assert np.all(pen[0] <= 0) and np.all(pen[1] >= 0) and np.all(pen[2] >= 0)   # signs of :993-995
    # greedy stochastic=False k=1 equals argmax decoding
    s2, _, _ = O.gen_sample(fi, fn, x, k=1, maxlen=8, stochastic=False, use_unk=True)
    s3, _, _ = O.gen_sample(fi, fn, x, k=1, maxlen=8, stochastic=True, argmax=True)
    assert list(s2[0]) == list(s3)
