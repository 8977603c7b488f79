import sys, io, contextlib, time, cProfile, pstats
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
run = lambda: nats.gen_sample_many(tparams, f_init, f_next, xs, opts, None, 10, 25, True, 1.0, 1.0, 1.0, concurrency=8, chunk=16)
run(); torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable(); t0 = time.time(); run(); torch.cuda.synchronize(); dt = time.time() - t0; pr.disable()
print('total %.1f ms for 32 sentences' % (dt * 1e3))
st = pstats.Stats(pr); st.sort_stats('tottime').print_stats(18)
