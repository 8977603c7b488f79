"""Where does a beam-search step spend host time? (cProfile over gen_sample: k=10, src_len=800, 3 penalties)"""
import sys, io, contextlib, cProfile, pstats, numpy as np, torch
sys.path.insert(0, '.')
import bench
from nats_b200 import nats
w = bench.WORKLOADS['c3']; opts = bench.options_of(w)
np.random.seed(1234); P = nats.init_params(opts)
P['ff_logit_b'][0] = -1e9
with contextlib.redirect_stdout(io.StringIO()): tp = nats.init_tparams(P)
f_init, f_next = nats.build_sampler(tp, opts, None)
rng = np.random.RandomState(4321)
x = np.array(rng.randint(2, w['n_words'], size=(800,)).tolist() + [0], dtype='int64').reshape(-1, 1)
nats.gen_sample(tp, f_init, f_next, x, opts, None, 10, 4, False, False, True, 1.0, 1.0, 1.0)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
nats.gen_sample(tp, f_init, f_next, x, opts, None, 10, 20, False, False, True, 1.0, 1.0, 1.0)
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(28); print(s.getvalue()[:6000])
