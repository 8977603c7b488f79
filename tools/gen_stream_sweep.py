"""sentences/s of gen_sample_many against the number of searches in flight (config 5 model, 32 sentences, 25 steps)"""
import sys, io, contextlib, time
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from nats_b200 import nats
w = bench.WORKLOADS['c5']; opts = bench.options_of(w)
np.random.seed(1234)
with contextlib.redirect_stdout(io.StringIO()):
    tparams = nats.init_tparams(nats.init_params(opts))
rng = np.random.RandomState(99)
xs = [np.array(rng.randint(2, w['n_words'], size=(800 - 7 * (i % 16),)).tolist() + [0], dtype='int64') for i in range(32)]
f_init, f_next = nats.build_sampler(tparams, opts, None)
b = tparams['ff_logit_b'].get_value(); b[0] = -1e9; tparams['ff_logit_b'].set_value(b)
for conc in (1, 2, 4, 6, 8, 12):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        nats.gen_sample_many(tparams, f_init, f_next, xs, opts, None, 10, 25, True, 1.0, 1.0, 1.0, concurrency=conc, chunk=16)
        torch.cuda.synchronize(); dt = time.time() - t0
    print('concurrency %2d: %.1f sentences/s (%.2f ms per sentence)' % (conc, len(xs) / dt, dt / len(xs) * 1e3))
