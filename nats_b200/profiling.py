"""Roofline probe used by bench.py: runs a few EAGER training steps with the library's per-kernel-class CUDA-event
timers switched on (nats_profile_*), and turns the result into the `roofline` object of the bench line."""
import ctypes
import json
import os

from . import _lib


def probe(graph, plan, f_update, steps=2, accum=None):
    """-> dict: per-class {ms, launches, GB/s, TFLOP/s} per step + 'launches_per_step'."""
    eng = graph.engine
    lib = eng.lib
    n = lib.nats_profile_num_classes()
    B = plan.shape[2]
    scale = 1.0 / (B * graph.world)
    torch = eng.torch

    def one_step():
        graph.enqueue_fwd(plan)
        graph.enqueue_bwd(plan, scale)
        if graph.world > 1:
            torch.distributed.all_reduce(graph.grads)
        graph.enqueue_clip()
        if accum is not None:
            accum()
        f_update(0.01)

    one_step()                                   # warm (eager path, timers off)
    torch.cuda.synchronize()
    _lib.check(lib.nats_profile_enable(eng.ctx, 1), 'nats_profile_enable')
    for _ in range(steps):
        one_step()
    ms = (ctypes.c_double * n)()
    fl = (ctypes.c_double * n)()
    by = (ctypes.c_double * n)()
    la = (ctypes.c_int64 * n)()
    _lib.check(lib.nats_profile_read(eng.ctx, n, ms, fl, by, la), 'nats_profile_read')
    lib.nats_profile_enable(eng.ctx, 0)
    out, total_ms, total_launch = {}, 0.0, 0
    for i in range(n):
        if la[i] == 0:
            continue
        name = lib.nats_profile_class_name(i).decode()
        t = ms[i] / steps
        out[name] = {'ms_per_step': t, 'launches_per_step': la[i] / steps, 'us_per_launch': 1e3 * ms[i] / la[i],
                     'algo_gbytes_per_step': by[i] / steps / 1e9, 'algo_gflop_per_step': fl[i] / steps / 1e9,
                     'GBps': (by[i] / 1e9) / (ms[i] / 1e3) if ms[i] > 0 else None,
                     'TFLOPps': (fl[i] / 1e12) / (ms[i] / 1e3) if ms[i] > 0 and fl[i] > 0 else None}
        total_ms += t
        total_launch += la[i] / steps
    for v in out.values():
        v['share'] = v['ms_per_step'] / total_ms if total_ms > 0 else None
    return {'classes': out, 'sum_kernel_ms_per_step': total_ms, 'launches_per_step': total_launch,
            'how': 'eager launches bracketed by CUDA events on the launch stream, %d steps' % steps}


def roofline_of(kernels, workload, peaks_path):
    """bench-line `roofline` object for the dominant kernel class (largest share of the step)."""
    peaks = {}
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            peaks = json.load(f)
    hbm = peaks.get('hbm_gbs', 6650.0)
    tf = peaks.get('bf16_tflops_sustained', 1400.0)
    which = 'measured' if peaks else 'fallback'
    cls = kernels['classes']
    name = max(cls, key=lambda k: cls[k]['ms_per_step'])
    k = cls[name]
    # DRAM bytes per launch of that kernel class from the committed `ncu --set full` capture (profiles/ncu_traffic.json)
    traffic, traffic_src = None, None
    tpath = os.path.join(os.path.dirname(peaks_path), 'profiles', 'ncu_traffic.json')
    if os.path.exists(tpath):
        with open(tpath) as f:
            t = json.load(f).get(name)
        if t:
            traffic, traffic_src = t.get('dram_bytes_per_launch'), t.get('source')
    if name.startswith('gemm'):
        ach = k['TFLOPps']
        return {'kernel': name, 'bound': 'tensor', 'achieved': ach, 'peak': tf, 'unit': 'TFLOP/s',
                'frac': ach / tf if ach else None, 'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': which + ' bf16 dense, sustained',
                'share_of_step': k['share'], 'us_per_launch': k['us_per_launch'],
                'note': 'v1 arithmetic is exact-fp32 FFMA (CUDA cores); the tensor-core ceiling is quoted as the '
                        'bound the GEMM-shaped work must be moved to'}
    ach = k['GBps']
    return {'kernel': name, 'bound': 'hbm', 'achieved': ach, 'peak': hbm, 'unit': 'GB/s',
            'frac': ach / hbm if ach else None, 'traffic': traffic, 'traffic_source': traffic_src,
            'algo_bytes_per_launch': (k['algo_gbytes_per_step'] * 1e9 / k['launches_per_step']) if k['launches_per_step'] else None,
            'peak_source': which + ' copy bandwidth',
            'share_of_step': k['share'], 'us_per_launch': k['us_per_launch']}
