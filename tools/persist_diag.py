import sys, ctypes, numpy as np, torch
sys.path.insert(0, '.')
import bench
from nats_b200 import nats, _lib
w = bench.WORKLOADS['c3']; opts = bench.options_of(w)
np.random.seed(1234); P = nats.init_params(opts)
import io, contextlib
with contextlib.redirect_stdout(io.StringIO()): tp = nats.init_tparams(P)
g = nats.build_model(tp, opts)[-1]
b = bench.make_batches(w, 1, 1)[0]
for _ in range(2): g.f_log_probs(*b)
torch.cuda.synchronize()
p = g.plan(w['Tx'], w['Ty'], w['B'])
lib = _lib.load()
# step_counters location: find via workspace view not exposed -> scan: use known offset by calling a tiny helper? read dbg through ctypes: we stored after bar (+16 uints)
# Locate step_counters by brute force: search ws for the counter values is fragile; instead re-run encoder only and time it
eng = g.engine
import time
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
Tx, Ty, B = p.shape
def enc():
    _lib.check(lib.nats_encoder_fwd(eng.ctx, eng.stream(), ctypes.byref(g.dims), ctypes.c_void_p(tp.flat.data_ptr()), ctypes.c_void_p(p.x.data_ptr()), ctypes.c_void_p(p.xm.data_ptr()), Tx, Ty, B, ctypes.c_void_p(p.ws.data_ptr()), p.ws_bytes))
enc(); torch.cuda.synchronize()
e0.record(); enc(); e1.record(); torch.cuda.synchronize()
print('encoder fwd total %.3f ms' % e0.elapsed_time(e1))
ptr = lib.nats_train_ws_view(ctypes.byref(g.dims), Tx, Ty, B, ctypes.c_void_p(p.ws.data_ptr()), b'step_counters')
off = (ptr - p.ws.data_ptr())
def show(name, lo):
    dbg = p.ws[off + 64: off + 64 + 64].view(torch.int64).cpu().numpy()
    d = dbg[lo:lo + 4]; tot = d.sum()
    print('%s CTA(0,0) cycles per step: product %.0f  reduce %.0f  gates %.0f  barrier %.0f  (total %.0f = %.2f us @1.965GHz)' % tuple([name] + list(d / Tx) + [tot / Tx, tot / Tx / 1965.0]))
show('fwd', 0)
g.grad_step(b[0], b[1], b[2], b[3], lambda *a, **k: None)
torch.cuda.synchronize()
show('bwd', 4)
