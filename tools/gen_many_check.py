"""Stress check of gen_sample_many at the config-5 model size: 96 sources of 20..800 words, beam 5, 30 steps, penalties on --
the interleaved searches (12 streams, batched encoders) must return what sentence-by-sentence gen_sample returns."""
import sys, io, contextlib, time
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from nats_b200 import nats
w = bench.WORKLOADS['c5']; opts = bench.options_of(w)
np.random.seed(1234)
with contextlib.redirect_stdout(io.StringIO()):
    tparams = nats.init_tparams(nats.init_params(opts))
rng = np.random.RandomState(5)
xs = [np.array(rng.randint(2, w['n_words'], size=(int(L),)).tolist() + [0], dtype='int64') for L in rng.randint(20, 801, size=96)]
f_init, f_next = nats.build_sampler(tparams, opts, None)
kw = dict(k=5, maxlen=30, use_unk=True, kl_factor=1.0, ctx_factor=1.0, state_factor=1.0)
t0 = time.time()
one = [nats.gen_sample(tparams, f_init, f_next, x[:, None], opts, stochastic=False, **kw) for x in xs]
torch.cuda.synchronize(); t1 = time.time()
bad = 0
for rep in range(3):
    many = nats.gen_sample_many(tparams, f_init, f_next, xs, opts, **kw)
    for i, ((s1, c1, a1), (s2, c2, a2)) in enumerate(zip(one, many)):
        same = [list(map(int, s)) for s in s1] == [list(map(int, s)) for s in s2]
        close = same and np.allclose(np.array(c1, 'float64'), np.array(c2, 'float64'), rtol=2e-4) and \
            all(np.allclose(np.array(h1), np.array(h2), rtol=2e-4, atol=1e-6) for h1, h2 in zip(a1, a2))
        if not close:
            bad += 1
            print('MISMATCH rep %d sentence %d (len %d): tokens same=%s' % (rep, i, len(xs[i]), same))
torch.cuda.synchronize(); t2 = time.time()
print('sequential %.2f s, 3 x many %.2f s, mismatches %d of %d' % (t1 - t0, t2 - t1, bad, 3 * len(xs)))
sys.exit(1 if bad else 0)
