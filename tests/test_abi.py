"""CPU-side checks of the drop-in boundary: the shared library loads, exports exactly the symbols that
include/nats_b200.h declares, the ctypes table covers them all, and the pure-host entry point
(nats_param_layout) reproduces the reference's parameter inventory (nats.py:613-654)."""
import os
import re
import subprocess

import numpy as np
import pytest

from nats_b200 import _lib
from oracle import nats_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module', autouse=True)
def built():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()


def _header_symbols():
    src = open(_lib.HEADER_PATH).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(nats_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    declared = _header_symbols()
    assert len(declared) >= 26
    out = subprocess.run(['nm', '-D', '--defined-only', _lib.LIB_PATH], capture_output=True, text=True, check=True)
    exported = sorted(set(l.split()[-1] for l in out.stdout.splitlines() if ' T ' in l and 'nats_' in l))
    assert exported == declared
    assert sorted(_lib.SIGNATURES.keys()) == declared
    lib = _lib.load()                       # dlopen + getattr of every symbol; no compute call (no GPU here)
    assert lib.nats_version() >= 100


def test_param_layout_matches_reference_inventory():
    opts = dict(n_words=50, dim_word=6, dim=8, dim_att=5)
    np.random.seed(0)
    P = O.init_params(dict(opts, encoder='gru', decoder='gru_cond'))
    views, total = _lib.param_layout(50, 6, 8, 5)
    assert [v[0] for v in views] == list(P.keys())
    used = np.zeros(total, dtype=bool)
    for (name, off, rows, cols, ld, ndim) in views:
        shape = (cols,) if ndim == 1 else (rows, cols)
        assert shape == P[name].shape, name
        assert off % 1 == 0 and ld >= cols
        idx = (off + np.arange(rows)[:, None] * ld + np.arange(cols)[None, :]).ravel()
        assert idx.max() < total
        assert not used[idx].any(), 'views overlap: ' + name
        used[idx] = True
    assert used.sum() == sum(v.size for v in P.values())
    # packed blocks: gate and candidate matrices side by side (one GEMM serves both)
    d = dict((v[0], v) for v in views)
    assert d['encoder_Ux'][1] == d['encoder_U'][1] + 2 * 8 and d['encoder_U'][4] == 24
    # C3 dims: 27,557,601 reference parameters
    views, total = _lib.param_layout(30000, 100, 1000, 100)
    assert sum(v[2] * v[3] for v in views) == 27557601 and total >= 27557601 and total % 32 == 0


def test_workspace_queries_are_host_only():
    import ctypes
    lib = _lib.load()
    d = _lib.Dims(30000, 100, 1000, 100)
    nb = lib.nats_train_workspace_bytes(ctypes.byref(d), 400, 30, 32)
    assert 1 << 28 < nb < 16 << 30
    assert lib.nats_sampler_workspace_bytes(ctypes.byref(d), 801, 10) > 0
    assert lib.nats_train_workspace_bytes(ctypes.byref(d), 0, 30, 32) < 0


def test_product_has_no_cpu_fallback():
    """Without a GPU the compiled callables must refuse, not silently compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from nats_b200 import nats
    np.random.seed(1)
    P = nats.init_params(dict(n_words=30, dim_word=4, dim=4, dim_att=3, encoder='gru', decoder='gru_cond'))
    with pytest.raises(_lib.NatsB200Error):
        nats.init_tparams(P)
    for f in os.listdir(os.path.join(ROOT, 'nats_b200')):
        if f.endswith('.py'):
            assert 'import oracle' not in open(os.path.join(ROOT, 'nats_b200', f)).read()
            assert 'from oracle' not in open(os.path.join(ROOT, 'nats_b200', f)).read()
