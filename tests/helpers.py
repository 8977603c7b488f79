"""Shared helpers for the test-suite (seeded toy problems for the oracle and the CUDA path)."""
import numpy as np

from oracle import nats_oracle as O


def toy_options(D=8, W=6, A=5, V=50):
    return dict(dim_word=W, dim=D, dim_att=A, n_words=V, encoder='gru', decoder='gru_cond')


def toy_params(opts, seed=1234, dtype='float64', bias_scale=0.1):
    """Reference init (nats.py:613-654) under numpy.random.seed(seed); biases/scale vectors are then
    perturbed so that every term of the graph is exercised (the reference initialises them to 0)."""
    np.random.seed(seed)
    P = O.init_params(opts)
    rng = np.random.RandomState(seed + 1)
    for k in P:
        if P[k].ndim == 1:
            P[k] = (bias_scale * rng.randn(*P[k].shape)).astype('float32')
        elif k.endswith(('_U_att', '_W_con', '_U_con', '_D_wei')):
            P[k] = (0.5 * rng.randn(*P[k].shape)).astype('float32')
        elif P[k].shape[0] != P[k].shape[1] or k == 'Wemb':
            P[k] = (0.3 * rng.randn(*P[k].shape)).astype('float32')
    return O.cast_params(P, dtype)


def ragged_batch(V, B=3, max_x=8, max_y=5, seed=7, min_x=2, min_y=1):
    rng = np.random.RandomState(seed)
    sx = [list(rng.randint(2, V, size=rng.randint(min_x, max_x + 1))) for _ in range(B)]
    sy = [list(rng.randint(2, V, size=rng.randint(min_y, max_y + 1))) for _ in range(B)]
    return O.prepare_data(sx, sy, n_words=V)


def full_batch(V, B, Tx, Ty, seed=1234):
    """BASELINE.md synthetic batch: fixed lengths so the padded shapes are exactly [Tx,B],[Ty,B]."""
    rng = np.random.RandomState(seed)
    sx = [list(rng.randint(2, V, size=Tx - 1)) for _ in range(B)]
    sy = [list(rng.randint(2, V, size=Ty - 1)) for _ in range(B)]
    return O.prepare_data(sx, sy, n_words=V)
