"""A short device-resident beam search (config 5, 6 steps) for ncu captures of the few-row kernels."""
import sys, io, contextlib
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from nats_b200 import nats
w = bench.WORKLOADS['c5']; opts = bench.options_of(w)
np.random.seed(1234)
with contextlib.redirect_stdout(io.StringIO()):
    tparams = nats.init_tparams(nats.init_params(opts))
rng = np.random.RandomState(4321)
x = np.array(rng.randint(2, w['n_words'], size=(800,)).tolist() + [0], dtype='int64').reshape(-1, 1)
f_init, f_next = nats.build_sampler(tparams, opts, None)
b = tparams['ff_logit_b'].get_value(); b[0] = -1e9; tparams['ff_logit_b'].set_value(b)
nats.gen_sample(tparams, f_init, f_next, x, opts, None, 10, int(sys.argv[1]) if len(sys.argv) > 1 else 6, False, False, True, 1.0, 1.0, 1.0)
torch.cuda.synchronize()
print('done')
