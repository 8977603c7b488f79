"""
Independent float64 restatement of the reference training graph (scripts/nats.py:658-772, 305-374,
454-609) in torch, used ONLY to cross-check the oracle's hand-derived backward through torch.autograd
(the role tensor.grad, nats.py:1340, plays in the reference).  Written separately from
oracle/nats_oracle.py on purpose (different structure: batch-major attention, fused gate matmuls), so a
shared mistake is unlikely.  Test infrastructure, CPU only.
"""
import torch


def _gru_seq(emb, mask, W, b, U, Wx, bx, Ux):
    T, B, _ = emb.shape
    D = Ux.shape[1]
    gates_in = emb @ W + b
    cand_in = emb @ Wx + bx
    h = emb.new_zeros(B, D)
    out = []
    for t in range(T):
        g = torch.sigmoid(h @ U + gates_in[t])
        r, u = g[:, :D], g[:, D:]
        cand = torch.tanh((h @ Ux) * r + cand_in[t])
        hn = u * h + (1 - u) * cand
        m = mask[t].unsqueeze(1)
        h = m * hn + (1 - m) * h
        out.append(h)
    return torch.stack(out)


def per_sample_cost(P, x, x_mask, y, y_mask):
    """P: dict name -> torch float64 tensor (requires_grad).  Returns cost [B]."""
    Tx, B = x.shape
    Ty = y.shape[0]
    D = P['encoder_Ux'].shape[1]
    xm, ym = x_mask, y_mask
    emb = P['Wemb'][x]
    embr = P['Wemb'][x.flip(0)]
    hf = _gru_seq(emb, xm, *[P['encoder_' + n] for n in ('W', 'b', 'U', 'Wx', 'bx', 'Ux')])
    hr = _gru_seq(embr, xm.flip(0), *[P['encoder_r_' + n] for n in ('W', 'b', 'U', 'Wx', 'bx', 'Ux')])
    ctx = torch.cat([hf, hr.flip(0)], dim=2)                       # [Tx,B,C]
    ctx_mean = (ctx * xm.unsqueeze(2)).sum(0) / xm.sum(0).unsqueeze(1)
    s = torch.tanh(ctx_mean @ P['ff_state_W'] + P['ff_state_b'])
    emby = P['Wemb'][y]
    embs = torch.cat([torch.zeros_like(emby[:1]), emby[:-1]], dim=0)
    d = lambda n: P['decoder_' + n]
    gates_in = embs @ d('W') + d('b')
    cand_in = embs @ d('Wx') + d('bx')
    ctxb = ctx.permute(1, 0, 2)                                    # [B,Tx,C]
    pctx = ctxb @ d('Wc_att') + d('b_att')                         # [B,Tx,A]
    xmb = xm.t()                                                   # [B,Tx]
    acc_c = ctx.new_zeros(B, 2 * D)
    acc_a = ctx.new_zeros(B, Tx)
    hs, cs = [], []
    for t in range(Ty):
        m = ym[t].unsqueeze(1)
        g = torch.sigmoid(s @ d('U') + gates_in[t])
        r, u = g[:, :D], g[:, D:]
        cand = torch.tanh((s @ d('Ux')) * r + cand_in[t])
        h1 = u * s + (1 - u) * cand
        h1 = m * h1 + (1 - m) * s
        att_h = torch.tanh(pctx + (h1 @ d('W_att')).unsqueeze(1) + acc_a.unsqueeze(2) * d('D_wei')[0])
        e = att_h @ d('U_att')[:, 0] + d('c_att')[0]               # [B,Tx]
        a = torch.exp(e) * xmb
        alpha = a / a.sum(1, keepdim=True)
        craw = torch.einsum('bt,btc->bc', alpha, ctxb)
        c = torch.tanh(d('U_con')[:, 0] * craw + acc_c * d('W_con')[:, 0])
        g2 = torch.sigmoid(h1 @ d('U_1') + d('b_1') + c @ d('W_1'))
        r2, u2 = g2[:, :D], g2[:, D:]
        cand2 = torch.tanh((h1 @ d('Ux_1') + d('bx_1')) * r2 + c @ d('Wx_1'))
        h2 = u2 * h1 + (1 - u2) * cand2
        h2 = m * h2 + (1 - m) * h1
        acc_c = acc_c + m * c
        acc_a = acc_a + m * alpha
        s = h2
        hs.append(h2); cs.append(c)
    Hs = torch.stack(hs); Cs = torch.stack(cs)
    L = torch.tanh(Hs @ P['ff_logit_lstm_W'] + P['ff_logit_lstm_b'] + embs @ P['ff_logit_prev_W']
                   + P['ff_logit_prev_b'] + Cs @ P['ff_logit_ctx_W'] + P['ff_logit_ctx_b'])
    logit = L @ P['ff_logit_W'] + P['ff_logit_b']
    logp = torch.log_softmax(logit, dim=2)
    nll = -logp.gather(2, y.unsqueeze(2)).squeeze(2)
    return (nll * ym).sum(0)
