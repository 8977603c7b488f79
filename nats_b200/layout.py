"""Host-side view of the packed flat parameter layout of libnats_b200 (nats_param_layout, include/nats_b200.h):
pack a reference-style dict of 43 arrays into one flat float32 vector and back.  Pure host code (no GPU)."""
from collections import OrderedDict

import numpy

from . import _lib


def dims_of(params):
    V, W = params['Wemb'].shape
    return int(V), int(W), int(params['encoder_Ux'].shape[0]), int(params['decoder_W_att'].shape[1])


def pack(params, tail=0):
    """dict (reference order / shapes) -> flat float32 [total + tail] in the device layout"""
    views, total = _lib.param_layout(*dims_of(params))
    flat = numpy.zeros(total + tail, dtype='float32')
    for (name, off, rows, cols, ld, ndim) in views:
        dst = numpy.lib.stride_tricks.as_strided(flat[off:], shape=(rows, cols), strides=(4 * ld, 4))
        dst[...] = numpy.asarray(params[name], dtype='float32').reshape(rows, cols)
    return flat


def unpack(flat, dims):
    """flat float32 -> OrderedDict name -> array (reference order / shapes)"""
    views, total = _lib.param_layout(*dims)
    out = OrderedDict()
    for (name, off, rows, cols, ld, ndim) in views:
        src = numpy.lib.stride_tricks.as_strided(flat[off:], shape=(rows, cols), strides=(4 * ld, 4))
        out[name] = numpy.array(src).reshape((cols,) if ndim == 1 else (rows, cols))
    return out
