"""Generate the golden fixtures under tests/golden/ from the float64 oracle (oracle/nats_oracle.py).

    python tests/golden/make_golden.py

The reference ships no golden vectors and cannot run here (Python 2 + Theano), so these pins are produced by
the CPU restatement ("parity unpinned", see the oracle header); the gradients they contain were cross-checked
against torch.autograd and central finite differences (tests/test_oracle_grad.py).  Contents:
  train_toy.npz    params (43), ragged batch, per-sample cost [B], all 43 gradients of mean cost, alphas, ctx
  adadelta_toy.npz parameters after one Adadelta step (clip_c = 1.0) on those gradients
  sampler_toy.npz  f_init outputs and 5 chained f_next outputs for one sentence
  beam_toy.npz     beam-search trace (k=3, lambda1=lambda2=lambda3=1.5): penalties, parents, words, costs
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import nats_oracle as O          # noqa: E402
from tests.helpers import toy_options, toy_params, ragged_batch   # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    opts = toy_options(D=8, W=6, A=5, V=50)
    P = toy_params(opts, seed=1234, dtype='float64')
    x, xm, y, ym = ragged_batch(opts['n_words'], B=3, max_x=8, max_y=5, seed=7)
    cost, cache = O.model_fwd(P, x, xm, y, ym)
    mean_cost, G, gnorm = O.f_grad(P, x, xm, y, ym)
    out = {'opt_dims': np.array([opts['n_words'], opts['dim_word'], opts['dim'], opts['dim_att']]),
           'x': x, 'x_mask': xm, 'y': y, 'y_mask': ym, 'cost': cost, 'mean_cost': mean_cost, 'gnorm': gnorm,
           'ctx': cache['ctx'], 'init_state': cache['init_state'], 'dec_h': cache['Hs'], 'dec_ctx': cache['Cs'],
           'dec_alpha': cache['As']}
    for k, v in P.items():
        out['p_' + k] = v
    for k, v in G.items():
        out['g_' + k] = v
    np.savez_compressed(os.path.join(HERE, 'train_toy.npz'), **out)

    _, Gc, _ = O.f_grad(P, x, xm, y, ym, clip_c=1.0)
    P1 = O.cast_params(P, 'float64')
    opt = O.Adadelta(P1)
    opt.grad_shared(Gc)
    opt.update(P1)
    np.savez_compressed(os.path.join(HERE, 'adadelta_toy.npz'), **{'p_' + k: v for k, v in P1.items()})

    xs = np.array([3, 17, 9, 4, 31, 5, 22, 0], 'int64')[:, None]
    init_state, ctx = O.f_init(P, xs)
    rec = {'x': xs, 'init_state': init_state, 'ctx': ctx}
    state, ac, aa = init_state, np.zeros((1, ctx.shape[2])), np.zeros((1, ctx.shape[0]))
    yprev = -np.ones((1,), 'int64')
    for t in range(5):
        probs, _, state, alT, c, ac, aa = O.f_next(P, yprev, ctx, state, ac, aa)
        rec.update({'probs%d' % t: probs, 'state%d' % t: state, 'alpha%d' % t: alT, 'ctxs%d' % t: c,
                    'acc_ctx%d' % t: ac, 'acc_alpha%d' % t: aa, 'yprev%d' % t: yprev})
        yprev = probs.argmax(1).astype('int64')
    np.savez_compressed(os.path.join(HERE, 'sampler_toy.npz'), **rec)

    P32 = O.cast_params(P, 'float32')
    fi = lambda x_: O.f_init(P32, x_)
    fn = lambda y_, ctx_, s_, ac_, aa_: O.f_next(P32, y_, ctx_, s_.astype('float32'), ac_.astype('float32'),
                                                 aa_.astype('float32'))
    tr = []
    samples, scores, _ = O.gen_sample(fi, fn, xs, k=3, maxlen=7, stochastic=False, use_unk=True, kl_factor=1.5,
                                      ctx_factor=1.5, state_factor=1.5, trace=tr)
    rec = {'x': xs, 'n_steps': len(tr), 'n_samples': len(samples), 'scores': np.array(scores, 'float32')}
    for i, s in enumerate(samples):
        rec['sample%d' % i] = np.array(s, 'int64')
    for t in tr:
        i = t['ii']
        rec['trans%d' % i] = t['trans']; rec['words%d' % i] = t['words']; rec['costs%d' % i] = t['costs']
        if t['pen'] is not None:
            rec['pen%d' % i] = t['pen']
    np.savez_compressed(os.path.join(HERE, 'beam_toy.npz'), **rec)
    print('golden fixtures written to', HERE)


if __name__ == '__main__':
    main()
