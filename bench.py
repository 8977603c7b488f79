#!/usr/bin/env python
"""bench.py -- decoder tokens/sec of the nats hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full training step of the hot path on one synthetic batch of the workload
(CNN/DM-shaped synthetic, BASELINE.json configs[2], the configuration the metric is quoted on):
    f_grad_shared(x, x_mask, y, y_mask)  = bi-GRU encoder, attention+distraction decoder scan, readout+softmax/NLL,
                                           hand-written backward, (N>1: ONE NCCL allreduce of the flat gradient),
                                           global-norm clip, Adadelta accumulators          (nats.py:1403)
    f_update(lrate)                      = Adadelta parameter update                          (nats.py:1409)
tokens = sum(y_mask) (= B*Ty per GPU), weak scaling (32 samples per GPU).

value : device-resident inputs, K steps replayed as CUDA graphs, CUDA-event timed, max over ranks.
e2e   : the same K steps through the reference-facing API with HOST numpy inputs: pinned staging + H2D of the
        batch and a D2H read of the cost inside the timed region, every step.
--impl reference : the CPU restatement of the reference (oracle/, float32, all host threads) on a bounded sample
        (8 of the 32 sentences) of the same workload -- the reference itself (Python 2 + Theano) cannot run here.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (src_len, tgt_len, dim, dim_word, dim_att, n_words, batch per GPU)
    'c3': dict(Tx=400, Ty=30, dim=1000, dim_word=100, dim_att=100, n_words=30000, B=32,
               name='CNN/DM-shaped synthetic: src_len=400, tgt_len=30, dim=1000, |V|=30000, batch=32/GPU'),
    'c2': dict(Tx=120, Ty=20, dim=500, dim_word=100, dim_att=100, n_words=4000, B=64,
               name='LCSTS-shaped synthetic: src_len=120, tgt_len=20, dim=500, |V|=4000, batch=64/GPU'),
}
METRIC = 'decoder tokens/sec (dim=1000, src=400, |V|=30k) @1/2/4/8 B200 vs Theano CPU'


def make_batches(w, n, seed, B=None):
    """BASELINE.md synthetic inputs: ids uniform in [2,V), fixed lengths -> padded shapes exactly [Tx,B],[Ty,B]."""
    from nats_b200.nats import prepare_data
    B = B or w['B']
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        sx = [list(rng.randint(2, w['n_words'], size=w['Tx'] - 1)) for _ in range(B)]
        sy = [list(rng.randint(2, w['n_words'], size=w['Ty'] - 1)) for _ in range(B)]
        out.append(prepare_data(sx, sy, n_words=w['n_words']))
    return out


def options_of(w):
    return dict(dim_word=w['dim_word'], dim=w['dim'], dim_att=w['dim_att'], n_words=w['n_words'], encoder='gru',
                decoder='gru_cond')


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for (_, r) in self.rows]
        for r in rows:
            f = [c.strip() for c in r.split(',')]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except Exception:
                continue
            for nme, val in zip(names, f[2:6]):
                if val.lower().startswith('active'):
                    reasons.add(nme)
        if not sm:
            return None
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(sm)}


def best_blas_threads():
    """OpenBLAS with every host thread (128 on the B200 boxes) is pathologically slow on the skinny products of the
    recurrence; pick the thread count that is fastest on representative shapes so the CPU baseline is a fair one."""
    try:
        from threadpoolctl import threadpool_limits
    except Exception:
        return None, os.cpu_count()
    ncpu = os.cpu_count() or 1
    rng = np.random.RandomState(0)
    a1, b1 = rng.randn(32, 1000).astype('float32'), rng.randn(1000, 3000).astype('float32')
    a2, b2 = rng.randn(960, 100).astype('float32'), rng.randn(100, 30000).astype('float32')
    best = (None, 1e30)
    for t in sorted(set([4, 8, 16, 32, 64, ncpu])):
        if t > ncpu:
            continue
        with threadpool_limits(limits=t):
            a1 @ b1; a2 @ b2
            t0 = time.perf_counter()
            for _ in range(20):
                a1 @ b1
            for _ in range(2):
                a2 @ b2
            dt = time.perf_counter() - t0
        if dt < best[1]:
            best = (t, dt)
    return threadpool_limits, best[0]


def cpu_reference_run(w, steps, warmup, sample_B=8):
    """Times the float32 CPU restatement on a bounded sample of the workload; returns tokens/s and details."""
    limiter, nthreads = best_blas_threads()
    if limiter is not None:
        with limiter(limits=nthreads):
            r = _cpu_reference_run(w, steps, warmup, sample_B)
    else:
        r = _cpu_reference_run(w, steps, warmup, sample_B)
    r['cores'] = nthreads
    r['sample'] += '; BLAS threads chosen by calibration out of %d host threads' % (os.cpu_count() or 1)
    return r


def _cpu_reference_run(w, steps, warmup, sample_B=8):
    from oracle import nats_oracle as O
    opts = options_of(w)
    np.random.seed(1234)
    P = O.init_params(opts)
    batches = make_batches(w, warmup + steps, seed=1234, B=sample_B)
    opt = O.Adadelta(P)
    times = []
    for i, (x, xm, y, ym) in enumerate(batches):
        t0 = time.perf_counter()
        cost, G, _ = O.f_grad(P, x, xm, y, ym, clip_c=100.)
        opt.grad_shared(G)
        opt.update(P)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    tokens = float(batches[0][3].sum())
    total = float(np.sum(times))
    return dict(value=tokens * len(times) / total, ms_per_step=1e3 * total / len(times), tokens_per_step=tokens,
                cores=os.cpu_count(), sample='%d of the %d sentences of each batch (Tx=%d, Ty=%d), %d timed steps'
                % (sample_B, w['B'], w['Tx'], w['Ty'], len(times)), cost=float(cost))


def extra_regions(nats, _lib, eng, graph, plan, tparams, opts, w, tokens_per_step):
    """SURVEY 8(d) side regions on one GPU (the headline `value` is R2, the train step):
    R1 decoder forward = gru_cond_layer scan + readout + NLL on a precomputed context (nats.py:737-770), device-timed;
    R3 beam step = gen_sample at beam 10, src_len 800, all three distraction penalties on, host-timed per step."""
    import ctypes
    import torch
    lib = eng.lib
    Tx, Ty, B = plan.shape
    vp = ctypes.c_void_p
    D = ctypes.byref(graph.dims)
    F, XM, Y, YM = vp(tparams.flat.data_ptr()), vp(plan.xm.data_ptr()), vp(plan.y.data_ptr()), vp(plan.ym.data_ptr())
    WS, C = vp(plan.ws.data_ptr()), vp(plan.cost.data_ptr())

    def r1():
        _lib.check(lib.nats_decoder_scan_fwd(eng.ctx, eng.stream(), D, F, Y, XM, YM, Tx, Ty, B, WS, plan.ws_bytes), 'scan')
        _lib.check(lib.nats_readout_nll_fwd(eng.ctx, eng.stream(), D, F, Y, YM, Tx, Ty, B, WS, plan.ws_bytes, C), 'readout')
    r1()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        r1()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    Dm, A, Cc = w['dim'], w['dim_att'], 2 * w['dim']
    step_bytes = 4.0 * (Tx * B * (Cc + A) + (12 * Dm * Dm + Dm * A) + B * (10 * Dm + 3 * Cc + 4 * Tx))
    peaks = {}
    if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')):
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peaks = json.load(f)
    hbm = peaks.get('hbm_gbs', 6650.0)
    out = {'R1_decoder_forward': {
        'value': tokens_per_step / (ms * 1e-3), 'unit': 'tokens/s', 'ms': ms, 'us_per_decoder_step': 1e3 * ms / Ty,
        'algo_bytes_per_step': step_bytes, 'achieved_GBps': step_bytes * Ty / (ms * 1e-3) / 1e9,
        'frac_of_hbm_peak': step_bytes * Ty / (ms * 1e-3) / 1e9 / hbm,
        'how': 'nats_decoder_scan_fwd + nats_readout_nll_fwd in one CUDA graph, CUDA events, %d replays' % n}}

    # R3: beam search, no hypothesis may finish (EOS logit pushed down), restored afterwards
    rng = np.random.RandomState(4321)
    xs = rng.randint(2, w['n_words'], size=(800,)).tolist() + [0]
    x = np.array(xs, dtype='int64').reshape(-1, 1)
    f_init, f_next = nats.build_sampler(tparams, opts, None)
    bsave = tparams['ff_logit_b'].get_value()
    bmod = bsave.copy()
    bmod[0] = -1e9
    tparams['ff_logit_b'].set_value(bmod)
    try:
        steps = 20
        nats.gen_sample(tparams, f_init, f_next, x, opts, None, 10, 4, False, False, True, 1.0, 1.0, 1.0)     # warm
        torch.cuda.synchronize()
        t0 = time.time()
        nats.gen_sample(tparams, f_init, f_next, x, opts, None, 10, steps, False, False, True, 1.0, 1.0, 1.0)
        torch.cuda.synchronize()
        dt = time.time() - t0
        t1 = time.time()
        f_init(x)
        torch.cuda.synchronize()
        t_init = time.time() - t1
    finally:
        tparams['ff_logit_b'].set_value(bsave)
    out['R3_beam_step'] = {'value': (dt - t_init) / steps * 1e3, 'unit': 'ms per beam step (k=10, src_len=800, 3 penalties)',
                           'f_init_ms': t_init * 1e3, 'steps': steps, 'hyp_tokens_per_s': 10 * steps / (dt - t_init),
                           'how': 'gen_sample wall clock incl. host bookkeeping, minus one f_init'}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c3', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--probe-steps', type=int, default=2, help='eager steps with per-kernel CUDA-event timing')
    ap.add_argument('--no-regions', action='store_true', help='skip the R1 (decoder forward) / R3 (beam step) side measurements')
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    warmup = max(args.warmup, 3) if args.impl == 'ours' else max(args.warmup, 1)

    config = {'workload': w['name'], 'region': 'train step: f_grad_shared + f_update (nats.py:1400-1411)',
              'optimizer': 'adadelta', 'clip_c': 100.0, 'global_batch': w['B'] * max(world, 1),
              'src_len': w['Tx'], 'tgt_len': w['Ty'], 'parallelism': 'dp%d' % max(world, 1),
              'l2_policy': 'per-step working set (saved activations + weights, > 2 GB) exceeds the 126 MB L2'}

    if args.impl == 'reference':
        if rank != 0:
            return 0
        r = cpu_reference_run(w, args.steps, warmup)
        line = {'impl': 'reference', 'metric': METRIC, 'value': r['value'], 'unit': 'tokens/s', 'n_gpus': args.gpus,
                'steps': args.steps, 'warmup': warmup, 'ms_per_step': r['ms_per_step'], 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config,
                'cpu_baseline': {'value': r['value'], 'unit': 'tokens/s', 'cores': r['cores'], 'kind': 'port',
                                 'sample': r['sample']},
                'e2e': {'value': r['value'], 'unit': 'tokens/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                'gpu_launches': 0,
                'note': 'reference = NumPy/OpenBLAS float32 restatement of scripts/nats.py (oracle/); Python 2 + '
                        'Theano are not installable here'}
        print(json.dumps(line))
        return 0

    os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')     # keep NCCL's banner off the one-JSON-line stdout
    import torch
    if world > 1:
        torch.cuda.set_device(local)
        torch.distributed.init_process_group('nccl', device_id=torch.device('cuda', local))
    from nats_b200 import nats, _lib
    eng = nats.get_engine()
    opts = options_of(w)
    np.random.seed(1234)
    params = nats.init_params(opts)
    import io
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        tparams = nats.init_tparams(params)
    graph = nats.build_model(tparams, opts)[-1].mean()
    graph.clip_c = 100.0
    f_grad_shared, f_update = nats.adadelta('lr', tparams, graph, None, graph)
    K = args.steps
    batches = make_batches(w, warmup + K, seed=1234 + rank)
    tokens_per_step = float(batches[0][3].sum())

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=eng.device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    # ---------------- e2e: host numpy in, host scalar out, every step ----------------
    costs = []
    for i in range(warmup):
        costs.append(float(f_grad_shared(*batches[i])))
        f_update(0.01)
    barrier()
    launches0 = eng.launches
    clk = ClockSampler(local) if rank == 0 else None
    t_wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(K):
        costs.append(float(f_grad_shared(*batches[warmup + i])))
        f_update(0.01)
    ev1.record()
    barrier()
    e2e_ms = max_over_ranks(ev0.elapsed_time(ev1))
    abi_calls = eng.launches - launches0
    plan = graph.plan(w['Tx'], w['Ty'], w['B'])
    h2d = plan.h2d_bytes()

    # ---------------- device-resident: same steps, inputs already in HBM ----------------
    def device_step():
        if world == 1:
            plan.graph_step.replay() if plan.graph_step is not None else None
        else:
            plan.graph_fb.replay()
            torch.distributed.all_reduce(graph.grads)
            plan.graph_post.replay()
        f_update(0.01)

    have_graphs = (plan.graph_step is not None) if world == 1 else (plan.graph_fb is not None)
    dev_ms = None
    if have_graphs:
        for _ in range(3):
            device_step()
        barrier()
        ev0.record()
        for _ in range(K):
            device_step()
        ev1.record()
        barrier()
        dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    t_wall1 = time.time()
    clocks = clk.stop(t_wall0, t_wall1) if clk is not None else None

    # ---------------- per-kernel probe (eager, CUDA events inside the library) ----------------
    roofline, kernels = None, None
    try:
        if args.probe_steps <= 0:
            raise RuntimeError('probe disabled')
        from nats_b200 import profiling
        kernels = profiling.probe(graph, plan, f_update, steps=args.probe_steps, accum=f_grad_shared.accum)
        roofline = profiling.roofline_of(kernels, w, os.path.join(ROOT, 'MEASURED_PEAKS.json'))
    except Exception as e:          # the probe is diagnostics; the headline numbers do not depend on it
        kernels = {'error': repr(e)}

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return 0

    regions = None
    if world == 1 and args.workload == 'c3' and not args.no_regions:
        try:
            regions = extra_regions(nats, _lib, eng, graph, plan, tparams, opts, w, tokens_per_step)
        except Exception as e:       # diagnostics only
            regions = {'error': repr(e)}

    total_tokens = tokens_per_step * max(world, 1) * K
    value_ms = dev_ms if dev_ms is not None else e2e_ms
    line = {
        'metric': METRIC, 'value': total_tokens / (value_ms * 1e-3), 'unit': 'tokens/s', 'n_gpus': max(world, 1),
        'steps': K, 'warmup': warmup, 'ms_per_step': value_ms / K, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config,
        'e2e': {'value': total_tokens / (e2e_ms * 1e-3), 'unit': 'tokens/s', 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': 4, 'ms_per_step': e2e_ms / K},
        'gpu_launches': None, 'abi_calls_timed': abi_calls,
        'clocks': clocks, 'roofline': roofline, 'regions': regions, 'kernels': kernels,
        'cost_first_last': [costs[0], costs[-1]],
    }
    if kernels and isinstance(kernels, dict) and 'launches_per_step' in kernels:
        line['gpu_launches'] = int(kernels['launches_per_step'] * K)
    if not args.no_cpu_baseline and world == 1:
        r = cpu_reference_run(w, steps=2, warmup=1)
        line['cpu_baseline'] = {'value': r['value'], 'unit': 'tokens/s', 'cores': r['cores'], 'kind': 'port',
                                'sample': r['sample'], 'ms_per_step': r['ms_per_step']}
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
