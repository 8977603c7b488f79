import sys, io, contextlib, time
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from nats_b200 import nats, _lib
w = bench.WORKLOADS['c5']; opts = bench.options_of(w)
np.random.seed(1234)
with contextlib.redirect_stdout(io.StringIO()):
    tparams = nats.init_tparams(nats.init_params(opts))
rng = np.random.RandomState(4321)
x = np.array(rng.randint(2, w['n_words'], size=(800,)).tolist() + [0], dtype='int64').reshape(-1, 1)
f_init, f_next = nats.build_sampler(tparams, opts, None)
b = tparams['ff_logit_b'].get_value(); b[0] = -1e9; tparams['ff_logit_b'].set_value(b)
lib = f_next.engine.lib
orig = lib.nats_beam_step
acc = [0.0, 0]
class Wrap:
    def __call__(self, *a):
        t = time.perf_counter(); r = orig(*a); acc[0] += time.perf_counter() - t; acc[1] += 1; return r
lib.nats_beam_step = Wrap()
for rep in range(3):
    acc[0] = 0.0; acc[1] = 0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    nats.gen_sample(tparams, f_init, f_next, x, opts, None, 10, 100, False, False, True, 1.0, 1.0, 1.0)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print('sentence %.2f ms; %d beam_step calls, host time inside the calls %.1f us each (%.2f ms total)' % (dt * 1e3, acc[1], acc[0] / max(acc[1], 1) * 1e6, acc[0] * 1e3))
