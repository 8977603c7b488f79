"""In-graph time of each phase of the training step (one CUDA graph per C-ABI piece, CUDA-event timed)."""
import sys, io, contextlib, ctypes, numpy as np, torch
sys.path.insert(0, '.')
import bench
from nats_b200 import nats, _lib
w = bench.WORKLOADS['c3']; opts = bench.options_of(w)
np.random.seed(1234); P = nats.init_params(opts)
with contextlib.redirect_stdout(io.StringIO()): tp = nats.init_tparams(P)
g = nats.build_model(tp, opts)[-1]
b = bench.make_batches(w, 1, 1)[0]
for _ in range(2): g.grad_step(b[0], b[1], b[2], b[3], lambda *a, **k: None)
torch.cuda.synchronize()
p = g.plan(w['Tx'], w['Ty'], w['B']); eng = g.engine; lib = eng.lib
Tx, Ty, B = p.shape
vp = ctypes.c_void_p
D = ctypes.byref(g.dims); F = vp(tp.flat.data_ptr()); X = vp(p.x.data_ptr()); XM = vp(p.xm.data_ptr()); Y = vp(p.y.data_ptr()); YM = vp(p.ym.data_ptr())
WS = vp(p.ws.data_ptr()); G = vp(g.grads.data_ptr()); C = vp(p.cost.data_ptr())
phases = [
    ('encoder_fwd', lambda: lib.nats_encoder_fwd(eng.ctx, eng.stream(), D, F, X, XM, Tx, Ty, B, WS, p.ws_bytes)),
    ('decoder_scan_fwd', lambda: lib.nats_decoder_scan_fwd(eng.ctx, eng.stream(), D, F, Y, XM, YM, Tx, Ty, B, WS, p.ws_bytes)),
    ('readout_nll_fwd', lambda: lib.nats_readout_nll_fwd(eng.ctx, eng.stream(), D, F, Y, YM, Tx, Ty, B, WS, p.ws_bytes, C)),
    ('readout_nll_bwd', lambda: lib.nats_readout_nll_bwd(eng.ctx, eng.stream(), D, F, Y, YM, Tx, Ty, B, WS, p.ws_bytes, ctypes.c_float(1.0 / B), G)),
    ('decoder_scan_bwd', lambda: lib.nats_decoder_scan_bwd(eng.ctx, eng.stream(), D, F, Y, XM, YM, Tx, Ty, B, WS, p.ws_bytes, G)),
    ('encoder_bwd', lambda: lib.nats_encoder_bwd(eng.ctx, eng.stream(), D, F, X, XM, Y, Tx, Ty, B, WS, p.ws_bytes, G)),
]
tot = 0.0
for name, fn in phases:
    _lib.check(fn(), name); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr): _lib.check(fn(), name)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): gr.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5; tot += ms
    print('%-18s %8.3f ms' % (name, ms), flush=True)
print('%-18s %8.3f ms (+ clip/optimizer outside)' % ('sum', tot))
