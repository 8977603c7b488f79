"""GPU busy time vs wall clock of the device-resident beam search (config 5): where does a beam step go?"""
import sys, time, re
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from nats_b200 import nats
w = bench.WORKLOADS['c5']; opts = bench.options_of(w)
np.random.seed(1234)
import io, contextlib
with contextlib.redirect_stdout(io.StringIO()):
    tparams = nats.init_tparams(nats.init_params(opts))
rng = np.random.RandomState(4321)
x = np.array(rng.randint(2, w['n_words'], size=(800,)).tolist() + [0], dtype='int64').reshape(-1, 1)
f_init, f_next = nats.build_sampler(tparams, opts, None)
b = tparams['ff_logit_b'].get_value(); b[0] = -1e9; tparams['ff_logit_b'].set_value(b)
steps = 40
nats.gen_sample(tparams, f_init, f_next, x, opts, None, 10, steps, False, False, True, 1.0, 1.0, 1.0)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
t0 = time.time()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    nats.gen_sample(tparams, f_init, f_next, x, opts, None, 10, steps, False, False, True, 1.0, 1.0, 1.0)
    torch.cuda.synchronize()
wall = time.time() - t0
evs = sorted((float(e.time_range.start), float(e.time_range.end), e.name) for e in prof.events() if 'cuda' in str(getattr(e, 'device_type', '')).lower())
busy, cover, per = 0.0, -1e30, {}
for st, en, name in evs:
    ex = max(0.0, en - max(st, cover)); cover = max(cover, en); busy += ex
    k = re.sub(r'\(anonymous namespace\)::|nats::|void ', '', name).split('(')[0][:56]
    d = per.setdefault(k, [0.0, 0]); d[0] += ex; d[1] += 1
print('wall %.2f ms (profiled), GPU busy %.2f ms, %d kernels, %.1f kernels/step, busy per step %.1f us' % (wall * 1e3, busy / 1e3, len(evs), len(evs) / steps, busy / steps))
for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:24]:
    print('  %-50s %8.1f us  %5d launches  %6.2f us/launch' % (k, v[0], v[1], v[0] / v[1]))
# timeline of one step in the middle: start offset, duration, gap to the previous end
sel = [i for i, e in enumerate(evs) if 'gather_rows' in e[2]]
if len(sel) > 22:
    lo, hi = sel[20], sel[21]
    t0 = evs[lo][0]; prev_end = t0
    print('--- timeline of one beam step (us): start  dur  gap_after_prev_end  name')
    for st, en, name in evs[lo:hi]:
        k = re.sub(r'\(anonymous namespace\)::|nats::|void ', '', name).split('(')[0][:44]
        print('  %8.1f %7.1f %7.1f  %s' % (st - t0, en - st, st - prev_end, k))
        prev_end = max(prev_end, en)
    print('  step span %.1f us' % (evs[hi][0] - t0))
