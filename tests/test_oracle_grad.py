"""The oracle's hand-derived backward == torch.autograd of an independent restatement == central
finite differences (float64).  This is what stands in for tensor.grad (nats.py:1340)."""
import numpy as np
import pytest
import torch

from oracle import nats_oracle as O
from tests.helpers import toy_options, toy_params, ragged_batch
from tests.torch_autograd_ref import per_sample_cost


@pytest.fixture(scope='module')
def problem():
    opts = toy_options()
    P = toy_params(opts)
    x, xm, y, ym = ragged_batch(opts['n_words'])
    return opts, P, x, xm, y, ym


def test_cost_matches_torch(problem):
    _, P, x, xm, y, ym = problem
    cost = O.f_log_probs(P, x, xm, y, ym)
    Pt = {k: torch.tensor(v, dtype=torch.float64) for k, v in P.items()}
    ct = per_sample_cost(Pt, torch.tensor(x), torch.tensor(xm, dtype=torch.float64), torch.tensor(y),
                         torch.tensor(ym, dtype=torch.float64)).numpy()
    np.testing.assert_allclose(cost, ct, rtol=1e-12, atol=1e-12)


def test_grads_match_autograd(problem):
    _, P, x, xm, y, ym = problem
    _, G, _ = O.f_grad(P, x, xm, y, ym)
    Pt = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in P.items()}
    ct = per_sample_cost(Pt, torch.tensor(x), torch.tensor(xm, dtype=torch.float64), torch.tensor(y),
                         torch.tensor(ym, dtype=torch.float64)).mean()
    ct.backward()
    assert list(G.keys()) == list(P.keys()) and len(G) == 43
    for k in P:
        g_ref = Pt[k].grad.numpy()
        # decoder_c_att's true gradient is 0 (softmax shift invariance) -> absolute floor
        assert np.abs(G[k] - g_ref).max() < 1e-9 * np.abs(g_ref).max() + 1e-13, k


def test_grads_match_finite_differences(problem):
    _, P, x, xm, y, ym = problem
    _, G, _ = O.f_grad(P, x, xm, y, ym)
    rng = np.random.RandomState(0)
    eps = 1e-6
    for k in P:
        for _ in range(3):
            idx = tuple(rng.randint(0, s) for s in P[k].shape)
            old = P[k][idx]
            P[k][idx] = old + eps
            cp = O.f_cost(P, x, xm, y, ym)
            P[k][idx] = old - eps
            cm = O.f_cost(P, x, xm, y, ym)
            P[k][idx] = old
            fd = (cp - cm) / (2 * eps)
            assert abs(fd - G[k][idx]) < 1e-6 * max(1., abs(fd)) + 1e-8, (k, idx, fd, G[k][idx])


def test_decay_and_clip(problem):
    _, P, x, xm, y, ym = problem
    c0, G0, n0 = O.f_grad(P, x, xm, y, ym)
    c1, G1, n1 = O.f_grad(P, x, xm, y, ym, decay_c=0.01)
    assert np.isclose(c1 - c0, 0.01 * sum((v ** 2).sum() for v in P.values()))
    for k in P:
        np.testing.assert_allclose(G1[k], G0[k] + 0.02 * P[k], rtol=1e-12, atol=1e-14)
    clip = 0.5 * n0
    _, G2, n2 = O.f_grad(P, x, xm, y, ym, clip_c=clip)
    assert np.isclose(n2, n0)
    assert np.isclose(np.sqrt(sum((g ** 2).sum() for g in G2.values())), clip)
    _, G3, _ = O.f_grad(P, x, xm, y, ym, clip_c=2 * n0)      # below threshold: untouched (:1350)
    for k in P:
        np.testing.assert_array_equal(G3[k], G0[k])
