"""Print the in-kernel phase timestamps of a few consecutive encoder steps (NATS_TRACE=<first skinny-GEMM launch number>).
Launch numbers count BN=32 products from context creation; pick a window inside the call that gets graph-captured
(third call) to see the replayed behaviour."""
import sys, io, contextlib, numpy as np, torch
sys.path.insert(0, '.')
import bench
from nats_b200 import nats
w = bench.WORKLOADS['c3']; opts = bench.options_of(w)
np.random.seed(1234); P = nats.init_params(opts)
with contextlib.redirect_stdout(io.StringIO()): tp = nats.init_tparams(P)
g = nats.build_model(tp, opts)[-1]
b = bench.make_batches(w, 1, 1)[0]
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    print('--- call', i + 1, flush=True)
    g.grad_step(b[0], b[1], b[2], b[3], lambda *a, **k: None)
    torch.cuda.synchronize()
