"""
oracle/nats_oracle.py -- CPU restatement (NumPy) of the hot path of lukecq1231/nats.

*** TEST INFRASTRUCTURE, NOT PRODUCT. ***  Only tests/, __graft_entry__.smoke() and the
`cpu_baseline` / `--impl reference` legs of bench.py may import this module.  The product
(nats_b200/) never imports it and has no CPU fallback.

*** PARITY UNPINNED. ***  The reference (scripts/nats.py) is a Python-2 / Theano program; neither
Python 2 nor Theano (un-vendored, un-pinned dependency, README.md:27-31; the API used --
theano.sandbox.rng_mrg, device=gpu0 -- is the 0.7-0.9 era) exists in this image and the reference
ships no tests, golden vectors or saved models.  This file therefore restates the *published
semantics* of the Theano ops at the reference's own call sites:

    tensor.dot                      -> matrix product
    tensor.nnet.sigmoid / tanh      -> 1/(1+exp(-x)) / tanh
    theano.scan(sequences, outputs_info, non_sequences) -> python loop carrying outputs_info
    tensor.nnet.softmax             -> row softmax
    tensor.nnet.categorical_crossentropy(p, idx) -> -log p[i, idx_i]
    tensor.grad                     -> reverse-mode derivative (hand-derived below, checked in
                                       tests/ by central finite differences AND against an independent
                                       torch.autograd float64 restatement)
    scipy.stats.entropy(p, q)       -> sum p^ log(p^/q^), p^,q^ normalised to 1   (checked against
    scipy.spatial.distance.cosine   -> 1 - u.v/(|u||v|)                            the real SciPy)

Every function cites the reference lines (scripts/nats.py unless noted) it follows.  All routines are
dtype-generic: they compute in the dtype of the parameters they are given (float64 = truth for the
parity tests; float32 = the "Theano CPU, floatX=float32" timing proxy, scripts/test.sh:3).
"""
from collections import OrderedDict
import copy

import numpy as np


# --------------------------------------------------------------------------------------------
# parameter initialisation                                                  nats.py:118-142, 251-260,
#                                                                           271-302, 378-451, 613-654
# --------------------------------------------------------------------------------------------
def ortho_weight(ndim):
    """nats.py:118-129 -- left singular vectors of a square Gaussian matrix (global numpy RNG)."""
    W = np.random.randn(ndim, ndim)
    u, _, _ = np.linalg.svd(W)
    return u.astype('float32')


def norm_weight(nin, nout=None, scale=0.01, ortho=True):
    """nats.py:132-142."""
    if nout is None:
        nout = nin
    if nout == nin and ortho:
        return ortho_weight(nin)
    return (scale * np.random.randn(nin, nout)).astype('float32')


def _init_ff(params, prefix, nin, nout, ortho=True):
    """nats.py:251-260."""
    params[prefix + '_W'] = norm_weight(nin, nout, scale=0.01, ortho=ortho)
    params[prefix + '_b'] = np.zeros((nout,), dtype='float32')


def _init_gru(params, prefix, nin, dim):
    """nats.py:271-302 (order of RNG draws preserved)."""
    params[prefix + '_W'] = np.concatenate([norm_weight(nin, dim), norm_weight(nin, dim)], axis=1)
    params[prefix + '_b'] = np.zeros((2 * dim,), dtype='float32')
    params[prefix + '_U'] = np.concatenate([ortho_weight(dim), ortho_weight(dim)], axis=1)
    params[prefix + '_Wx'] = norm_weight(nin, dim)
    params[prefix + '_bx'] = np.zeros((dim,), dtype='float32')
    params[prefix + '_Ux'] = ortho_weight(dim)


def _init_gru_cond(params, prefix, nin, dim, dimctx, dimatt):
    """nats.py:378-451 (order of keys and of RNG draws preserved)."""
    params[prefix + '_W'] = np.concatenate([norm_weight(nin, dim), norm_weight(nin, dim)], axis=1)
    params[prefix + '_U'] = np.concatenate([ortho_weight(dim), ortho_weight(dim)], axis=1)
    params[prefix + '_b'] = np.zeros((2 * dim,), dtype='float32')
    params[prefix + '_Wx'] = norm_weight(nin, dim)
    params[prefix + '_Ux'] = ortho_weight(dim)
    params[prefix + '_bx'] = np.zeros((dim,), dtype='float32')
    params[prefix + '_U_1'] = np.concatenate([ortho_weight(dim), ortho_weight(dim)], axis=1)
    params[prefix + '_W_1'] = norm_weight(dimctx, dim * 2)
    params[prefix + '_b_1'] = np.zeros((2 * dim,), dtype='float32')
    params[prefix + '_Wx_1'] = norm_weight(dimctx, dim)
    params[prefix + '_Ux_1'] = ortho_weight(dim)
    params[prefix + '_bx_1'] = np.zeros((dim,), dtype='float32')
    params[prefix + '_W_att'] = norm_weight(dim, dimatt)
    params[prefix + '_Wc_att'] = norm_weight(dimctx, dimatt)
    params[prefix + '_b_att'] = np.zeros((dimatt,), dtype='float32')
    params[prefix + '_U_att'] = norm_weight(dimatt, 1)
    params[prefix + '_c_att'] = np.zeros((1,), dtype='float32')
    params[prefix + '_W_con'] = norm_weight(dimctx, 1)
    params[prefix + '_U_con'] = norm_weight(dimctx, 1)
    params[prefix + '_D_wei'] = norm_weight(1, dimatt)


def init_params(options):
    """nats.py:613-654 -- the 43 tensors, in the reference's OrderedDict order."""
    params = OrderedDict()
    W, D, A, V = options['dim_word'], options['dim'], options['dim_att'], options['n_words']
    params['Wemb'] = norm_weight(V, W)
    _init_gru(params, 'encoder', W, D)
    _init_gru(params, 'encoder_r', W, D)
    _init_ff(params, 'ff_state', 2 * D, D)
    _init_gru_cond(params, 'decoder', W, D, 2 * D, A)
    _init_ff(params, 'ff_logit_lstm', D, W, ortho=False)
    _init_ff(params, 'ff_logit_prev', W, W, ortho=False)
    _init_ff(params, 'ff_logit_ctx', 2 * D, W, ortho=False)
    _init_ff(params, 'ff_logit', W, V)
    return params


def cast_params(params, dtype):
    return OrderedDict((k, np.asarray(v, dtype=dtype)) for k, v in params.items())


# --------------------------------------------------------------------------------------------
# batch layout                                                                     nats.py:200-247
# --------------------------------------------------------------------------------------------
def prepare_data(seqs_x, seqs_y, maxlen=None, n_words=30000):
    """nats.py:200-247: truncate to maxlen-1, zero-pad, masks have len+1 ones (the EOS row)."""
    lengths_x = [len(s) for s in seqs_x]
    lengths_y = [len(s) for s in seqs_y]
    if maxlen is not None:
        seqs_x = [s[:maxlen - 1] if len(s) >= maxlen else s for s in seqs_x]
        seqs_y = [s[:maxlen - 1] if len(s) >= maxlen else s for s in seqs_y]
        lengths_x = [len(s) for s in seqs_x]
        lengths_y = [len(s) for s in seqs_y]
        if len(lengths_x) < 1 or len(lengths_y) < 1:
            return None, None, None, None
    n = len(seqs_x)
    Tx = int(np.max(lengths_x)) + 1
    Ty = int(np.max(lengths_y)) + 1
    x = np.zeros((Tx, n), dtype='int64')
    y = np.zeros((Ty, n), dtype='int64')
    xm = np.zeros((Tx, n), dtype='float32')
    ym = np.zeros((Ty, n), dtype='float32')
    for i, (sx, sy) in enumerate(zip(seqs_x, seqs_y)):
        x[:lengths_x[i], i] = sx
        xm[:lengths_x[i] + 1, i] = 1.
        y[:lengths_y[i], i] = sy
        ym[:lengths_y[i] + 1, i] = 1.
    return x, xm, y, ym


# --------------------------------------------------------------------------------------------
# layers, forward
# --------------------------------------------------------------------------------------------
def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def gru_layer_fwd(P, prefix, emb, mask):
    """nats.py:305-374.  emb [T,B,W], mask [T,B] -> H [T,B,D] and the per-step cache."""
    W, b, U = P[prefix + '_W'], P[prefix + '_b'], P[prefix + '_U']
    Wx, bx, Ux = P[prefix + '_Wx'], P[prefix + '_bx'], P[prefix + '_Ux']
    T, B, _ = emb.shape
    D = Ux.shape[1]
    dt = U.dtype
    xg = emb @ W + b          # :328-329
    xc = emb @ Wx + bx        # :331-332
    H = np.zeros((T, B, D), dt)
    R = np.zeros((T, B, D), dt); Ug = np.zeros((T, B, D), dt)
    Cn = np.zeros((T, B, D), dt); Pp = np.zeros((T, B, D), dt)
    h = np.zeros((B, D), dt)  # :360
    for t in range(T):
        pre = h @ U + xg[t]                  # :337-338
        r = _sigmoid(pre[:, :D])             # :341
        u = _sigmoid(pre[:, D:])             # :342
        p = h @ Ux                           # :345
        c = np.tanh(p * r + xc[t])           # :346-350
        hn = u * h + (1. - u) * c            # :353
        m = mask[t][:, None].astype(dt)
        h = m * hn + (1. - m) * h            # :354
        H[t] = h; R[t] = r; Ug[t] = u; Cn[t] = c; Pp[t] = p
    return H, dict(emb=emb, mask=mask, H=H, R=R, U=Ug, C=Cn, P=Pp)


def gru_layer_bwd(P, prefix, cache, dH, G):
    """Reverse of gru_layer_fwd (what tensor.grad, nats.py:1340, derives through the scan :365-372).
    dH [T,B,D] = gradient arriving at every output state.  Accumulates parameter grads into G,
    returns d emb [T,B,W]."""
    W, U, Wx, Ux = P[prefix + '_W'], P[prefix + '_U'], P[prefix + '_Wx'], P[prefix + '_Ux']
    emb, mask, H = cache['emb'], cache['mask'], cache['H']
    T, B, D = H.shape
    dt = U.dtype
    dxg = np.zeros((T, B, 2 * D), dt)
    dxc = np.zeros((T, B, D), dt)
    dU = np.zeros_like(U); dUx = np.zeros_like(Ux)
    dh_next = np.zeros((B, D), dt)
    for t in range(T - 1, -1, -1):
        hp = H[t - 1] if t > 0 else np.zeros((B, D), dt)
        r, u, c, p = cache['R'][t], cache['U'][t], cache['C'][t], cache['P'][t]
        m = mask[t][:, None].astype(dt)
        dh = dH[t] + dh_next
        dhn = m * dh
        dhp = (1. - m) * dh + dhn * u
        du = dhn * (hp - c)
        dc = dhn * (1. - u)
        dpc = dc * (1. - c * c)
        dp = dpc * r
        dr = dpc * p
        dg = np.concatenate([dr * r * (1. - r), du * u * (1. - u)], axis=1)
        dxg[t] = dg
        dxc[t] = dpc
        dhp = dhp + dg @ U.T + dp @ Ux.T
        dU += hp.T @ dg
        dUx += hp.T @ dp
        dh_next = dhp
    W_ = emb.shape[2]
    e2 = emb.reshape(T * B, W_)
    G[prefix + '_W'] += e2.T @ dxg.reshape(T * B, 2 * D)
    G[prefix + '_b'] += dxg.sum((0, 1))
    G[prefix + '_Wx'] += e2.T @ dxc.reshape(T * B, D)
    G[prefix + '_bx'] += dxc.sum((0, 1))
    G[prefix + '_U'] += dU
    G[prefix + '_Ux'] += dUx
    return dxg @ W.T + dxc @ Wx.T


def gru_cond_step(P, m_, x_, xx_, h_, acc_ctx, acc_alpha, pctx_, cc_, context_mask, prefix='decoder'):
    """nats.py:498-572 (_step_slice), one decoder step.
    m_ [B], x_ [B,2D], xx_ [B,D], h_ [B,D], acc_ctx [B,C], acc_alpha [B,Tx], pctx_ [Tx,B,A], cc_ [Tx,B,C],
    context_mask [Tx,B] or None.  Returns (h2, ctx, alpha.T, acc_ctx', acc_alpha'), cache."""
    p = prefix
    U, Ux = P[p + '_U'], P[p + '_Ux']
    W_1, U_1, b_1 = P[p + '_W_1'], P[p + '_U_1'], P[p + '_b_1']
    Wx_1, Ux_1, bx_1 = P[p + '_Wx_1'], P[p + '_Ux_1'], P[p + '_bx_1']
    W_att, U_att, c_att = P[p + '_W_att'], P[p + '_U_att'], P[p + '_c_att']
    W_con, U_con, D_wei = P[p + '_W_con'], P[p + '_U_con'], P[p + '_D_wei']
    D = Ux.shape[1]
    dt = U.dtype
    m = m_[:, None].astype(dt)

    # GRU_2 (:505-518) -- note bx lives in xx_, outside the reset gate
    g1 = _sigmoid(h_ @ U + x_)
    r1, u1 = g1[:, :D], g1[:, D:]
    p1 = h_ @ Ux
    c1 = np.tanh(p1 * r1 + xx_)
    h1 = u1 * h_ + (1. - u1) * c1
    h1 = m * h1 + (1. - m) * h_

    # attention with distraction over past attention weights (:527-541)
    ps = h1 @ W_att                                                     # [B,A]
    z = np.tanh(pctx_ + ps[None, :, :] + acc_alpha.T[:, :, None] * D_wei[0][None, None, :])   # :528-533
    e = z @ U_att[:, 0] + c_att[0]                                      # [Tx,B]  :535-536
    a = np.exp(e)                                                       # :537 (no max-subtraction)
    if context_mask is not None:
        a = a * context_mask.astype(dt)                                 # :538-539
    alpha = a / a.sum(0, keepdims=True)                                 # :540
    craw = np.einsum('tb,tbc->bc', alpha, cc_)                          # :541  [B,C] (no [Tx,B,C] temporary)

    # distraction over past context vectors (:545-546); U_con, W_con are per-channel scales
    ctx = np.tanh(U_con[:, 0][None, :] * craw + acc_ctx * W_con[:, 0][None, :])

    # GRU_1 (:551-565) -- note bx_1 is inside the reset gate
    g2 = _sigmoid(h1 @ U_1 + b_1 + ctx @ W_1)
    r2, u2 = g2[:, :D], g2[:, D:]
    p2 = h1 @ Ux_1 + bx_1
    c2 = np.tanh(p2 * r2 + ctx @ Wx_1)
    h2 = u2 * h1 + (1. - u2) * c2
    h2 = m * h2 + (1. - m) * h1

    # accumulators (:569-570)
    acc_ctx_n = m * ctx + acc_ctx
    acc_alpha_n = m * alpha.T + acc_alpha

    cache = dict(m=m, h_=h_, r1=r1, u1=u1, p1=p1, c1=c1, h1=h1, ps=ps, z=z, alpha=alpha, craw=craw,
                 ctx=ctx, acc_ctx=acc_ctx, acc_alpha=acc_alpha, r2=r2, u2=u2, p2=p2, c2=c2)
    return (h2, ctx, alpha.T, acc_ctx_n, acc_alpha_n), cache


def gru_cond_layer_fwd(P, emb, mask, context, context_mask, init_state, prefix='decoder'):
    """nats.py:454-609 with one_step=False: scan over Ty steps."""
    p = prefix
    Ty, B, _ = emb.shape
    Tx, _, C = context.shape
    D = P[p + '_Ux'].shape[1]
    dt = P[p + '_U'].dtype
    xx = emb @ P[p + '_Wx'] + P[p + '_bx']                  # :487-488
    xg = emb @ P[p + '_W'] + P[p + '_b']                    # :490-491
    pctx = context @ P[p + '_Wc_att'] + P[p + '_b_att']     # :493-494
    h = init_state
    acc_ctx = np.zeros((B, C), dt)                          # :599-603
    acc_alpha = np.zeros((B, Tx), dt)
    Hs = np.zeros((Ty, B, D), dt); Cs = np.zeros((Ty, B, C), dt); As = np.zeros((Ty, B, Tx), dt)
    caches = []
    for t in range(Ty):
        (h, ctx, alT, acc_ctx, acc_alpha), ch = gru_cond_step(
            P, mask[t], xg[t], xx[t], h, acc_ctx, acc_alpha, pctx, context, context_mask, prefix)
        Hs[t] = h; Cs[t] = ctx; As[t] = alT
        caches.append(ch)
    cache = dict(emb=emb, mask=mask, context=context, context_mask=context_mask, init_state=init_state,
                 pctx=pctx, steps=caches)
    return (Hs, Cs, As, acc_ctx, acc_alpha), cache


def gru_cond_layer_bwd(P, cache, dHs, dCs, G, prefix='decoder'):
    """Reverse of gru_cond_layer_fwd.  dHs [Ty,B,D], dCs [Ty,B,C] = gradients arriving from the readout
    at every step's h2 / ctx outputs.  Returns (d emb [Ty,B,W], d context [Tx,B,C], d init_state [B,D]).
    Carries d acc_ctx [B,C] and d acc_alpha [B,Tx] backwards (suffix sums through :569-570)."""
    p = prefix
    U, Ux = P[p + '_U'], P[p + '_Ux']
    W_1, U_1 = P[p + '_W_1'], P[p + '_U_1']
    Wx_1, Ux_1 = P[p + '_Wx_1'], P[p + '_Ux_1']
    W_att, U_att = P[p + '_W_att'], P[p + '_U_att']
    W_con, U_con, D_wei = P[p + '_W_con'], P[p + '_U_con'], P[p + '_D_wei']
    Wc_att = P[p + '_Wc_att']
    emb, cc, pctx = cache['emb'], cache['context'], cache['pctx']
    Ty, B, Wd = emb.shape
    Tx, _, C = cc.shape
    D = Ux.shape[1]
    dt = U.dtype
    ucon, wcon, dwei, uatt = U_con[:, 0], W_con[:, 0], D_wei[0], U_att[:, 0]

    dxg = np.zeros((Ty, B, 2 * D), dt); dxx = np.zeros((Ty, B, D), dt)
    dpctx = np.zeros_like(pctx)
    dcraws = np.zeros((Ty, B, C), dt)
    dh = np.zeros((B, D), dt)
    dacc_ctx = np.zeros((B, C), dt); dacc_alpha = np.zeros((B, Tx), dt)
    for t in range(Ty - 1, -1, -1):
        s = cache['steps'][t]
        m, h_, h1, ctx = s['m'], s['h_'], s['h1'], s['ctx']
        dh2 = dHs[t] + dh
        dctx = dCs[t] + m * dacc_ctx                  # acc_ctx' = m ctx + acc_ctx  (:569)
        dalphaT = m * dacc_alpha                      # acc_alpha' = m alpha.T + acc_alpha (:570)
        # GRU_1 backward (:551-565)
        r2, u2, p2, c2 = s['r2'], s['u2'], s['p2'], s['c2']
        dhn2 = m * dh2
        dh1 = (1. - m) * dh2 + dhn2 * u2
        du2 = dhn2 * (h1 - c2)
        dc2 = dhn2 * (1. - u2)
        dpc2 = dc2 * (1. - c2 * c2)
        dp2 = dpc2 * r2
        dr2 = dpc2 * p2
        dg2 = np.concatenate([dr2 * r2 * (1. - r2), du2 * u2 * (1. - u2)], axis=1)
        dh1 = dh1 + dg2 @ U_1.T + dp2 @ Ux_1.T
        dctx = dctx + dg2 @ W_1.T + dpc2 @ Wx_1.T
        G[p + '_U_1'] += h1.T @ dg2; G[p + '_Ux_1'] += h1.T @ dp2
        G[p + '_W_1'] += ctx.T @ dg2; G[p + '_Wx_1'] += ctx.T @ dpc2
        G[p + '_b_1'] += dg2.sum(0); G[p + '_bx_1'] += dp2.sum(0)
        # ctx distraction backward (:545-546)
        dq = dctx * (1. - ctx * ctx)
        G[p + '_U_con'][:, 0] += (dq * s['craw']).sum(0)
        G[p + '_W_con'][:, 0] += (dq * s['acc_ctx']).sum(0)
        dcraw = dq * ucon[None, :]
        dacc_ctx = dacc_ctx + dq * wcon[None, :]
        # attention backward (:527-541)
        alpha, z = s['alpha'], s['z']                                     # [Tx,B], [Tx,B,A]
        dalpha = np.einsum('tbc,bc->tb', cc, dcraw) + dalphaT.T           # :541
        dcraws[t] = dcraw                                                 # d cc summed after the loop
        de = alpha * (dalpha - (alpha * dalpha).sum(0, keepdims=True))    # :537-540 (masked softmax)
        G[p + '_c_att'][0] += de.sum()
        G[p + '_U_att'][:, 0] += np.einsum('tb,tba->a', de, z)
        dzp = de[:, :, None] * uatt[None, None, :] * (1. - z * z)        # [Tx,B,A]
        dpctx += dzp
        dps = dzp.sum(0)                                                  # [B,A]
        dacc_alpha = dacc_alpha + (dzp @ dwei).T                          # :532
        G[p + '_D_wei'][0] += np.einsum('bt,tba->a', s['acc_alpha'], dzp)
        dh1 = dh1 + dps @ W_att.T
        G[p + '_W_att'] += h1.T @ dps
        # GRU_2 backward (:505-518)
        r1, u1, p1, c1 = s['r1'], s['u1'], s['p1'], s['c1']
        dhn1 = m * dh1
        dh_ = (1. - m) * dh1 + dhn1 * u1
        du1 = dhn1 * (h_ - c1)
        dc1 = dhn1 * (1. - u1)
        dpc1 = dc1 * (1. - c1 * c1)
        dp1 = dpc1 * r1
        dr1 = dpc1 * p1
        dg1 = np.concatenate([dr1 * r1 * (1. - r1), du1 * u1 * (1. - u1)], axis=1)
        dxg[t] = dg1
        dxx[t] = dpc1
        dh_ = dh_ + dg1 @ U.T + dp1 @ Ux.T
        G[p + '_U'] += h_.T @ dg1; G[p + '_Ux'] += h_.T @ dp1
        dh = dh_
    e2 = emb.reshape(Ty * B, Wd)
    G[p + '_W'] += e2.T @ dxg.reshape(Ty * B, 2 * D)
    G[p + '_b'] += dxg.sum((0, 1))
    G[p + '_Wx'] += e2.T @ dxx.reshape(Ty * B, D)
    G[p + '_bx'] += dxx.sum((0, 1))
    G[p + '_Wc_att'] += cc.reshape(Tx * B, C).T @ dpctx.reshape(Tx * B, -1)
    G[p + '_b_att'] += dpctx.sum((0, 1))
    alphas = np.stack([s_['alpha'] for s_ in cache['steps']])            # [Ty,Tx,B]
    # d cc[t,b,:] = sum_s alpha_s[t,b] dcraw_s[b,:]  (one batched product instead of Ty rank-1 updates)
    dcc = np.matmul(alphas.transpose(2, 1, 0), dcraws.transpose(1, 0, 2)).transpose(1, 0, 2)
    dcc = dcc + dpctx @ Wc_att.T
    demb = dxg @ P[p + '_W'].T + dxx @ P[p + '_Wx'].T
    return demb, dcc, dh


# --------------------------------------------------------------------------------------------
# training graph                                                                  nats.py:658-772
# --------------------------------------------------------------------------------------------
def model_fwd(P, x, x_mask, y, y_mask):
    """build_model, nats.py:658-772.  Returns per-sample cost [B] and the cache for model_bwd."""
    Wemb = P['Wemb']
    dt = Wemb.dtype
    Tx, B = x.shape
    Ty = y.shape[0]
    Wd = Wemb.shape[1]
    xm = x_mask.astype(dt); ym = y_mask.astype(dt)
    xr = x[::-1]; xrm = xm[::-1]                                      # :692-693
    emb = Wemb[x.flatten()].reshape(Tx, B, Wd)                        # :700-701
    embr = Wemb[xr.flatten()].reshape(Tx, B, Wd)                      # :706-707
    Hf, cf = gru_layer_fwd(P, 'encoder', emb, xm)                     # :702-704
    Hr, cr = gru_layer_fwd(P, 'encoder_r', embr, xrm)                 # :708-710
    ctx = np.concatenate([Hf, Hr[::-1]], axis=2)                      # :713
    xsum = xm.sum(0)
    ctx_mean = (ctx * xm[:, :, None]).sum(0) / xsum[:, None]          # :717
    init_state = np.tanh(ctx_mean @ P['ff_state_W'] + P['ff_state_b'])   # :723-724
    emby = Wemb[y.flatten()].reshape(Ty, B, Wd)                       # :730-731
    embs = np.zeros_like(emby); embs[1:] = emby[:-1]                  # :732-734
    (Hs, Cs, As, _, _), cd = gru_cond_layer_fwd(P, embs, ym, ctx, xm, init_state)   # :737-742
    pre = (Hs @ P['ff_logit_lstm_W'] + P['ff_logit_lstm_b'] + embs @ P['ff_logit_prev_W']
           + P['ff_logit_prev_b'] + Cs @ P['ff_logit_ctx_W'] + P['ff_logit_ctx_b'])     # :753-758
    L = np.tanh(pre)                                                  # :759
    logit = L @ P['ff_logit_W'] + P['ff_logit_b']                     # :760-761
    lg = logit.reshape(Ty * B, -1)
    mx = lg.max(1, keepdims=True)
    lse = mx[:, 0] + np.log(np.exp(lg - mx).sum(1))                   # softmax :763 (stable form)
    yf = y.flatten()                                                  # :767
    nll = lse - lg[np.arange(Ty * B), yf]                             # :768
    cost = (nll.reshape(Ty, B) * ym).sum(0)                           # :769-770
    cache = dict(x=x, xr=xr, xm=xm, ym=ym, y=y, cf=cf, cr=cr, ctx=ctx, xsum=xsum, ctx_mean=ctx_mean,
                 init_state=init_state, embs=embs, cd=cd, Hs=Hs, Cs=Cs, As=As, L=L, lg=lg, lse=lse)
    return cost, cache


def zero_grads(P):
    return OrderedDict((k, np.zeros_like(v)) for k, v in P.items())


def model_bwd(P, cache, dcost):
    """Gradient of sum_b dcost[b] * cost[b] w.r.t. all 43 tensors (reference: tensor.grad, :1340)."""
    G = zero_grads(P)
    dt = P['Wemb'].dtype
    x, xr, y = cache['x'], cache['xr'], cache['y']
    xm, ym = cache['xm'], cache['ym']
    Tx, B = x.shape
    Ty = y.shape[0]
    D = P['encoder_Ux'].shape[1]
    Hs, Cs, L, embs = cache['Hs'], cache['Cs'], cache['L'], cache['embs']
    Wd = embs.shape[2]
    # softmax + CE backward (:763-770)
    probs = np.exp(cache['lg'] - cache['lse'][:, None])
    dlg = probs
    dlg[np.arange(Ty * B), y.flatten()] -= 1.
    dlg *= (ym * np.asarray(dcost, dt)[None, :]).reshape(Ty * B, 1)
    L2 = L.reshape(Ty * B, Wd)
    G['ff_logit_W'] += L2.T @ dlg
    G['ff_logit_b'] += dlg.sum(0)
    dpre = (dlg @ P['ff_logit_W'].T) * (1. - L2 * L2)
    H2 = Hs.reshape(Ty * B, D); C2 = Cs.reshape(Ty * B, 2 * D); E2 = embs.reshape(Ty * B, Wd)
    G['ff_logit_lstm_W'] += H2.T @ dpre; G['ff_logit_lstm_b'] += dpre.sum(0)
    G['ff_logit_prev_W'] += E2.T @ dpre; G['ff_logit_prev_b'] += dpre.sum(0)
    G['ff_logit_ctx_W'] += C2.T @ dpre; G['ff_logit_ctx_b'] += dpre.sum(0)
    dHs = (dpre @ P['ff_logit_lstm_W'].T).reshape(Ty, B, D)
    dCs = (dpre @ P['ff_logit_ctx_W'].T).reshape(Ty, B, 2 * D)
    dembs = (dpre @ P['ff_logit_prev_W'].T).reshape(Ty, B, Wd)
    demb_d, dctx, dinit = gru_cond_layer_bwd(P, cache['cd'], dHs, dCs, G)
    dembs = dembs + demb_d
    # shifted target embedding (:730-734): embs[t] = Wemb[y[t-1]], embs[0] = 0
    np.add.at(G['Wemb'], y[:-1].flatten(), dembs[1:].reshape((Ty - 1) * B, Wd))
    # init state (:717-724)
    dpre0 = dinit * (1. - cache['init_state'] ** 2)
    G['ff_state_W'] += cache['ctx_mean'].T @ dpre0
    G['ff_state_b'] += dpre0.sum(0)
    dmean = dpre0 @ P['ff_state_W'].T
    dctx = dctx + xm[:, :, None] * (dmean / cache['xsum'][:, None])[None, :, :]
    # encoders (:702-713)
    demb_f = gru_layer_bwd(P, 'encoder', cache['cf'], dctx[:, :, :D], G)
    demb_r = gru_layer_bwd(P, 'encoder_r', cache['cr'], dctx[::-1, :, D:], G)
    np.add.at(G['Wemb'], x.flatten(), demb_f.reshape(Tx * B, Wd))
    np.add.at(G['Wemb'], xr.flatten(), demb_r.reshape(Tx * B, Wd))
    return G


def f_log_probs(P, x, x_mask, y, y_mask):
    """nats.py:1320."""
    return model_fwd(P, x, x_mask, y, y_mask)[0]


def f_cost(P, x, x_mask, y, y_mask, decay_c=0.):
    """nats.py:1323-1336: mean over the batch, plus L2."""
    cost = model_fwd(P, x, x_mask, y, y_mask)[0].mean()
    if decay_c > 0.:
        cost = cost + decay_c * sum((v ** 2).sum() for v in P.values())
    return cost


def f_grad(P, x, x_mask, y, y_mask, decay_c=0., clip_c=-1.):
    """nats.py:1323-1353: (mean cost [+L2], clipped gradients in parameter order, unclipped norm)."""
    cost_b, cache = model_fwd(P, x, x_mask, y, y_mask)
    B = cost_b.shape[0]
    G = model_bwd(P, cache, np.full((B,), 1.0 / B))
    cost = cost_b.mean()
    if decay_c > 0.:
        cost = cost + decay_c * sum((v ** 2).sum() for v in P.values())
        for k in G:
            G[k] += 2. * decay_c * P[k]
    g2 = sum((g ** 2).sum() for g in G.values())
    if clip_c > 0. and g2 > clip_c ** 2:                              # :1344-1353
        sc = clip_c / np.sqrt(g2)
        for k in G:
            G[k] = G[k] * sc
    return cost, G, np.sqrt(g2)


# --------------------------------------------------------------------------------------------
# optimisers                                                            nats.py:1106-1206
# --------------------------------------------------------------------------------------------
class Adadelta:
    """nats.py:1145-1173.  grad_shared(): zg <- g, rg2 <- rho rg2 + (1-rho) g^2 (:1156-1158);
    update(): ud = -sqrt(ru2+eps)/sqrt(rg2+eps) zg; ru2 <- rho ru2 + (1-rho) ud^2; p <- p+ud (:1163-1168)."""

    def __init__(self, P, epsilon=1e-6, rho=0.95):
        self.eps, self.rho = epsilon, rho
        self.zg = zero_grads(P); self.ru2 = zero_grads(P); self.rg2 = zero_grads(P)

    def grad_shared(self, G):
        for k, g in G.items():
            self.zg[k] = g.copy()
            self.rg2[k] = self.rho * self.rg2[k] + (1 - self.rho) * g ** 2

    def update(self, P, lr=None):
        for k in P:
            ud = -np.sqrt(self.ru2[k] + self.eps) / np.sqrt(self.rg2[k] + self.eps) * self.zg[k]
            self.ru2[k] = self.rho * self.ru2[k] + (1 - self.rho) * ud ** 2
            P[k] = P[k] + ud


class Adam:
    """nats.py:1106-1142 (the reference's "1-beta" parametrisation: b1=0.1, b2=0.001)."""

    def __init__(self, P):
        self.lr0, self.b1, self.b2, self.e = 0.0002, 0.1, 0.001, 1e-8
        self.g = zero_grads(P); self.m = zero_grads(P); self.v = zero_grads(P)
        self.i = 0.

    def grad_shared(self, G):
        for k, g in G.items():
            self.g[k] = g.copy()

    def update(self, P, lr=None):
        i_t = self.i + 1.
        fix1 = 1. - self.b1 ** i_t
        fix2 = 1. - self.b2 ** i_t
        lr_t = self.lr0 * (np.sqrt(fix2) / fix1)
        for k in P:
            g = self.g[k]
            m_t = self.b1 * g + (1. - self.b1) * self.m[k]
            v_t = self.b2 * g ** 2 + (1. - self.b2) * self.v[k]
            P[k] = P[k] - lr_t * (m_t / (np.sqrt(v_t) + self.e))
            self.m[k], self.v[k] = m_t, v_t
        self.i = i_t


class RMSprop:
    """nats.py:1176-1206 (Graves form with momentum 0.9)."""

    def __init__(self, P):
        self.zg = zero_grads(P); self.rg = zero_grads(P); self.rg2 = zero_grads(P); self.ud = zero_grads(P)

    def grad_shared(self, G):
        for k, g in G.items():
            self.zg[k] = g.copy()
            self.rg[k] = 0.95 * self.rg[k] + 0.05 * g
            self.rg2[k] = 0.95 * self.rg2[k] + 0.05 * g ** 2

    def update(self, P, lr=None):
        for k in P:
            ud = 0.9 * self.ud[k] - 1e-4 * self.zg[k] / np.sqrt(self.rg2[k] - self.rg[k] ** 2 + 1e-4)
            self.ud[k] = ud
            P[k] = P[k] + ud


# --------------------------------------------------------------------------------------------
# sampler graph                                                                   nats.py:776-874
# --------------------------------------------------------------------------------------------
def f_init(P, x):
    """nats.py:789-817: unmasked bi-GRU, ctx.mean(0), ff_state.  x int64 [Tx,n] -> (init_state, ctx)."""
    Wemb = P['Wemb']
    dt = Wemb.dtype
    Tx, n = x.shape
    Wd = Wemb.shape[1]
    ones = np.ones((Tx, n), dt)                                       # mask=None -> all valid (:317-318)
    xr = x[::-1]
    emb = Wemb[x.flatten()].reshape(Tx, n, Wd)
    embr = Wemb[xr.flatten()].reshape(Tx, n, Wd)
    Hf, _ = gru_layer_fwd(P, 'encoder', emb, ones)
    Hr, _ = gru_layer_fwd(P, 'encoder_r', embr, ones)
    ctx = np.concatenate([Hf, Hr[::-1]], axis=2)                      # :807
    ctx_mean = ctx.mean(0)                                            # :810
    init_state = np.tanh(ctx_mean @ P['ff_state_W'] + P['ff_state_b'])
    return init_state, ctx


def f_next(P, y, ctx, state, acc_ctx, acc_alpha, rng=None):
    """nats.py:821-871.  y int64 [n] (-1 = BOS -> zero embedding, :827-829).
    Returns [probs, sample, state', alpha.T, ctxs, acc_ctx', acc_alpha'] in the reference order (:870)."""
    Wemb = P['Wemb']
    dt = Wemb.dtype
    n = y.shape[0]
    emb = np.where((y < 0)[:, None], np.zeros((1, Wemb.shape[1]), dt), Wemb[np.maximum(y, 0)])
    xx = emb @ P['decoder_Wx'] + P['decoder_bx']
    xg = emb @ P['decoder_W'] + P['decoder_b']
    pctx = ctx @ P['decoder_Wc_att'] + P['decoder_b_att']
    ones = np.ones((n,), dt)                                          # mask=None (:472-473)
    (h2, c, alT, acc_ctx_n, acc_alpha_n), _ = gru_cond_step(
        P, ones, xg, xx, state, acc_ctx, acc_alpha, pctx, ctx, None)
    pre = (h2 @ P['ff_logit_lstm_W'] + P['ff_logit_lstm_b'] + emb @ P['ff_logit_prev_W']
           + P['ff_logit_prev_b'] + c @ P['ff_logit_ctx_W'] + P['ff_logit_ctx_b'])
    logit = np.tanh(pre) @ P['ff_logit_W'] + P['ff_logit_b']
    mx = logit.max(1, keepdims=True)
    ex = np.exp(logit - mx)
    probs = ex / ex.sum(1, keepdims=True)                             # :861
    if rng is None:
        sample = probs.argmax(1)
    else:                                                             # :864 (Theano MRG stream: unpinned)
        sample = np.array([rng.multinomial(1, pr.astype('float64') / pr.astype('float64').sum()).argmax()
                           for pr in probs])
    return [probs, sample.astype('int64'), h2, alT, c, acc_ctx_n, acc_alpha_n]


# --------------------------------------------------------------------------------------------
# beam search with distraction re-ranking                                       nats.py:879-1076
# --------------------------------------------------------------------------------------------
def entropy(pk, qk):
    """scipy.stats.entropy(pk, qk): KL(pk||qk) after normalising both to sum 1 (call site :990)."""
    pk = np.asarray(pk); qk = np.asarray(qk)
    pk = pk / pk.sum()
    qk = qk / qk.sum()
    with np.errstate(divide='ignore', invalid='ignore'):
        t = np.where(pk > 0, pk * np.log(pk / qk), np.where(pk == 0, 0., np.inf))
    return t.sum()


def cosine(u, v):
    """scipy.spatial.distance.cosine(u, v) = 1 - u.v/(|u||v|)  (call sites :991-992)."""
    u = np.asarray(u); v = np.asarray(v)
    return 1. - np.dot(u, v) / (np.sqrt(np.dot(u, u)) * np.sqrt(np.dot(v, v)))


def distraction_scores(hist_alphas, hist_ctxs, hist_states, dec_alphas, ctxs, next_state,
                       kl_factor, ctx_factor, state_factor):
    """nats.py:982-995 for all live hypotheses: (-l1*min KL, +l2*max cos-dist, +l3*max cos-dist)."""
    live_k = len(hist_alphas)
    out = np.zeros((3, live_k), dtype='float32')
    for idx in range(live_k):
        a_reg = [entropy(al, dec_alphas[idx, :]) for al in hist_alphas[idx]]
        c_reg = [cosine(c_, ctxs[idx, :]) for c_ in hist_ctxs[idx]]
        s_reg = [cosine(s_, next_state[idx, :]) for s_ in hist_states[idx]]
        out[0, idx] = -kl_factor * min(a_reg)
        out[1, idx] = ctx_factor * max(c_reg)
        out[2, idx] = state_factor * max(s_reg)
    return out


def gen_sample(f_init_, f_next_, x, k=1, maxlen=30, stochastic=True, argmax=False, use_unk=False,
               kl_factor=0, ctx_factor=0, state_factor=0, trace=None):
    """nats.py:879-1076, literal py3 restatement.  f_init_(x), f_next_(y, ctx, state, acc_ctx, acc_alpha)
    are callables with the reference signatures.  `trace` (a list) receives per-step debug records."""
    if k > 1:
        assert not stochastic, 'Beam search does not support stochastic sampling'
    sample, sample_score, sample_dec_alphas = [], [], []
    if stochastic:
        sample_score = 0
    live_k, dead_k = 1, 0
    hyp_samples = [[]] * live_k
    hyp_scores = np.zeros(live_k).astype('float32')
    hyp_dec_alphas = [[]] * live_k
    hyp_ctxs = [[]] * live_k
    hyp_states_dis = [[]] * live_k

    next_state, ctx0 = f_init_(x)                                       # :951-952
    next_w = -1 * np.ones((1,)).astype('int64')
    acc_ctx = np.zeros((live_k, ctx0.shape[2])).astype('float32')
    acc_alpha = np.zeros((live_k, ctx0.shape[0])).astype('float32')

    for ii in range(maxlen):
        ctx = np.tile(ctx0, [live_k, 1])                                # :958
        ret = f_next_(next_w, ctx, next_state, acc_ctx, acc_alpha)
        next_p, next_w, next_state, dec_alphas, ctxs, acc_ctx, acc_alpha = ret
        if stochastic:
            nw = next_p[0].argmax() if argmax else next_w[0]
            sample.append(nw)
            sample_score += next_p[0, nw]
            if nw == 0:
                break
            continue
        if not use_unk:
            next_p[:, 1] = 1e-20                                        # :974
        cand_scores = hyp_scores[:, None] - np.log(next_p)              # :976
        cand_flat = cand_scores.flatten()
        ranks_flat = cand_flat.argsort()[:(k - dead_k)]
        pen = None
        if ii > 0 and (kl_factor > 0. or ctx_factor > 0. or state_factor > 0.):   # :981
            pen = distraction_scores(hyp_dec_alphas, hyp_ctxs, hyp_states_dis, dec_alphas, ctxs, next_state,
                                     kl_factor, ctx_factor, state_factor)
            new_cand = cand_scores + pen[0][:, None] + pen[1][:, None] + pen[2][:, None]   # :997
            ranks_flat = new_cand.flatten().argsort()[:(k - dead_k)]
        voc_size = next_p.shape[1]
        trans_indices = ranks_flat // voc_size                          # :1002 (py2 int division)
        word_indices = ranks_flat % voc_size
        costs = cand_flat[ranks_flat]                                   # :1004 (un-penalised cost kept)
        if trace is not None:
            trace.append(dict(ii=ii, live_k=live_k, pen=None if pen is None else pen.copy(),
                              trans=trans_indices.copy(), words=word_indices.copy(), costs=costs.copy()))

        new_samples, new_scores = [], np.zeros(k - dead_k).astype('float32')
        new_states, new_alphas, new_ctxs, new_acc_ctx, new_acc_alpha, new_sdis = [], [], [], [], [], []
        for idx, (ti, wi) in enumerate(zip(trans_indices, word_indices)):
            new_samples.append(hyp_samples[ti] + [wi])
            new_scores[idx] = copy.copy(costs[idx])
            new_states.append(copy.copy(next_state[ti]))
            new_alphas.append(hyp_dec_alphas[ti] + [copy.copy(dec_alphas[ti, :])])
            new_ctxs.append(hyp_ctxs[ti] + [copy.copy(ctxs[ti, :])])
            new_acc_ctx.append(copy.copy(acc_ctx[ti]))
            new_acc_alpha.append(copy.copy(acc_alpha[ti]))
            new_sdis.append(hyp_states_dis[ti] + [copy.copy(next_state[ti, :])])

        new_live_k = 0
        hyp_samples, hyp_scores, hyp_states = [], [], []
        hyp_dec_alphas, hyp_ctxs, hyp_acc_alpha, hyp_acc_ctx, hyp_states_dis = [], [], [], [], []
        for idx in range(len(new_samples)):
            if new_samples[idx][-1] == 0:                               # :1037
                sample.append(new_samples[idx])
                sample_score.append(new_scores[idx])
                sample_dec_alphas.append(new_alphas[idx])
                dead_k += 1
            else:
                new_live_k += 1
                hyp_samples.append(new_samples[idx]); hyp_scores.append(new_scores[idx])
                hyp_states.append(new_states[idx]); hyp_dec_alphas.append(new_alphas[idx])
                hyp_ctxs.append(new_ctxs[idx]); hyp_acc_ctx.append(new_acc_ctx[idx])
                hyp_acc_alpha.append(new_acc_alpha[idx]); hyp_states_dis.append(new_sdis[idx])
        hyp_scores = np.array(hyp_scores)
        live_k = new_live_k
        if new_live_k < 1 or dead_k >= k:
            break
        next_w = np.array([w[-1] for w in hyp_samples])
        next_state = np.array(hyp_states)
        acc_ctx = np.array(hyp_acc_ctx)
        acc_alpha = np.array(hyp_acc_alpha)

    if not stochastic and live_k > 0:                                   # :1068-1074
        for idx in range(live_k):
            sample.append(hyp_samples[idx])
            sample_score.append(hyp_scores[idx])
            sample_dec_alphas.append(hyp_dec_alphas[idx])
    return sample, sample_score, sample_dec_alphas
