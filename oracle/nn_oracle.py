"""Oracle for the AFF (CvT) and NEG (BiGRU) forward passes (SURVEY.md §8 rows a2.*).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Floating-point kernel, so the
oracle is a plain torch CPU restatement (fp32 by default, fp64 on request) that
consumes a ``state_dict`` in the reference's own key layout.

Reference: clairs/model.py (cited as M:line); hyper-parameters are derived from
tensor shapes, never hard-coded (clairs/predict.py:513-553 uses two configs).
"""

from __future__ import annotations

import math

import torch
import torch.nn.functional as F

DIM_HEAD = 64          # M:103 (Attention default, never overridden)
LN_EPS = 1e-5          # M:58
BN_EPS = 1e-5          # nn.BatchNorm2d default, M:96


def _channel_ln(x, g, b):
    """M:57-67: normalise over channels with POPULATION std, eps added to std."""
    mean = x.mean(dim=1, keepdim=True)
    std = x.var(dim=1, unbiased=False, keepdim=True).sqrt()
    return (x - mean) / (std + LN_EPS) * g.reshape(1, -1, 1) + b.reshape(1, -1, 1)


def _mid(w):
    """3x3 kernels act on H=1 maps: only the middle row is ever used."""
    return w[:, :, 1, :]


def _dw_bn_pw(x, sd, p, stride):
    """M:91-100 DepthWiseConv2d: depth-wise 3-tap -> eval BatchNorm -> 1x1, no biases."""
    c = x.shape[1]
    y = F.conv1d(x, _mid(sd[p + '.net.0.weight']), stride=stride, padding=1, groups=c)
    scale = sd[p + '.net.1.weight'] / torch.sqrt(sd[p + '.net.1.running_var'] + BN_EPS)
    shift = sd[p + '.net.1.bias'] - sd[p + '.net.1.running_mean'] * scale
    y = y * scale.reshape(1, -1, 1) + shift.reshape(1, -1, 1)
    return F.conv1d(y, sd[p + '.net.2.weight'][:, :, 0, :])


def _attention(x, sd, p):
    """M:102-132."""
    b, c, w = x.shape
    q = _dw_bn_pw(x, sd, p + '.to_q', 1)
    kv = _dw_bn_pw(x, sd, p + '.to_kv', 2)
    inner = q.shape[1]
    heads = inner // DIM_HEAD
    k, v = kv[:, :inner], kv[:, inner:]

    def split(t):      # 'b (h d) w -> (b h) w d'
        return t.reshape(b, heads, DIM_HEAD, t.shape[-1]).permute(0, 1, 3, 2)

    q, k, v = split(q), split(k), split(v)
    dots = torch.matmul(q, k.transpose(-1, -2)) * (DIM_HEAD ** -0.5)
    attn = torch.softmax(dots, dim=-1)
    out = torch.matmul(attn, v)                                  # [b, h, w, d]
    out = out.permute(0, 1, 3, 2).reshape(b, inner, w)           # 'b (h d) w'
    return F.conv1d(out, sd[p + '.to_out.0.weight'][:, :, 0, :], sd[p + '.to_out.0.bias'])


def _feed_forward(x, sd, p):
    """M:78-89: 1x1 -> exact-erf GELU -> 1x1."""
    h = F.conv1d(x, sd[p + '.net.0.weight'][:, :, 0, :], sd[p + '.net.0.bias'])
    h = F.gelu(h)
    return F.conv1d(h, sd[p + '.net.3.weight'][:, :, 0, :], sd[p + '.net.3.bias'])


def _stage(x, sd, name):
    """M:194-198: embed conv (3-tap, stride 2, pad 1) -> channel LN -> transformer (M:143-147)."""
    x = F.conv1d(x, _mid(sd[name + '.0.weight']), sd[name + '.0.bias'], stride=2, padding=1)
    x = _channel_ln(x, sd[name + '.1.g'], sd[name + '.1.b'])
    depth = 0
    while '%s.2.layers.%d.0.norm.g' % (name, depth) in sd:
        depth += 1
    for d in range(depth):
        p = '%s.2.layers.%d' % (name, d)
        x = _attention(_channel_ln(x, sd[p + '.0.norm.g'], sd[p + '.0.norm.b']), sd, p + '.0.fn') + x
        x = _feed_forward(_channel_ln(x, sd[p + '.1.norm.g'], sd[p + '.1.norm.b']), sd, p + '.1.fn') + x
    return x


def head_names(sd, negational):
    """4 heads (SNV) or 6 (indel): M:214-224 / 332-345 / 420-433 / 503-520."""
    names = ['a', 'c', 'g', 't', 'i', 'd']
    if negational:
        names = ['n' + n for n in names]
    return [n for n in names if n + '_fc2.weight' in sd]


def _heads(feat, sd, negational):
    """M:239-253 (and twins): fc1 -> SELU -> per head fc2 -> SELU -> fc3 -> SELU."""
    h = F.selu(F.linear(feat, sd['fc1.weight'], sd['fc1.bias']))
    outs = []
    for n in head_names(sd, negational):
        y = F.selu(F.linear(h, sd[n + '_fc2.weight'], sd[n + '_fc2.bias']))
        outs.append(F.selu(F.linear(y, sd[n + '_fc3.weight'], sd[n + '_fc3.bias'])))
    return torch.stack(outs, dim=1)          # [B, H, 2]


def _cast(sd, dtype):
    return {k: v.to(dtype) for k, v in sd.items() if torch.is_floating_point(v)}


def aff_forward(x, sd, dtype=torch.float32):
    """CvT / CvT_Indel forward, M:231-261 / 348-384.  x: [B,33,34] -> logits [B,H,2] (post-SELU)."""
    sd = _cast(sd, dtype)
    x = torch.as_tensor(x, dtype=dtype).permute(0, 2, 1)      # [B, 34 channels, 33 positions]
    with torch.no_grad():
        for name in ('layer1', 'layer2', 'layer3'):
            if name + '.0.weight' in sd:
                x = _stage(x, sd, name)
        feat = x.reshape(x.shape[0], -1)                      # channel-major flatten: c*W + w
        return _heads(feat, sd, negational=False)


def _gru_direction(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """torch.nn.GRU cell, gate order r|z|n; n = tanh(W_in x + b_in + r*(W_hn h + b_hn))."""
    b, t, _ = x.shape
    hidden = w_hh.shape[1]
    gi = torch.matmul(x, w_ih.t()) + b_ih                     # [B, T, 3H]
    h = x.new_zeros(b, hidden)
    out = x.new_empty(b, t, hidden)
    steps = range(t - 1, -1, -1) if reverse else range(t)
    for s in steps:
        gh = torch.matmul(h, w_hh.t()) + b_hh
        r = torch.sigmoid(gi[:, s, :hidden] + gh[:, :hidden])
        z = torch.sigmoid(gi[:, s, hidden:2 * hidden] + gh[:, hidden:2 * hidden])
        n = torch.tanh(gi[:, s, 2 * hidden:] + r * gh[:, 2 * hidden:])
        h = (1.0 - z) * n + z * h
        out[:, s] = h
    return out


def _bigru(x, sd, name):
    fwd = _gru_direction(x, sd[name + '.weight_ih_l0'], sd[name + '.weight_hh_l0'],
                         sd[name + '.bias_ih_l0'], sd[name + '.bias_hh_l0'], reverse=False)
    bwd = _gru_direction(x, sd[name + '.weight_ih_l0_reverse'], sd[name + '.weight_hh_l0_reverse'],
                         sd[name + '.bias_ih_l0_reverse'], sd[name + '.bias_hh_l0_reverse'], reverse=True)
    return torch.cat([fwd, bwd], dim=-1)


def neg_forward(x, sd, dtype=torch.float32):
    """BiGRU_NACGT / _Indel forward, M:440-467 / 527-560.  x: [B,33,34] -> logits [B,H,2]."""
    sd = _cast(sd, dtype)
    x = torch.as_tensor(x, dtype=dtype)
    with torch.no_grad():
        y = _bigru(x, sd, 'lstm')           # attribute names say lstm; the modules are nn.GRU (M:412-417)
        y = _bigru(y, sd, 'lstm_2')
        feat = y.reshape(y.shape[0], -1)    # time-major flatten: t*384 + j
        return _heads(feat, sd, negational=True)


def softmax_heads(logits):
    """clairs/predict.py:574, 659-684: Softmax(dim=1) on each [B,2] head."""
    return torch.softmax(torch.as_tensor(logits), dim=-1)


# ----------------------------------------------------------------------------------------------
# deterministic synthetic weights in the reference's state_dict layout (no checkpoints offline): the
# generator lives in the product package (bench.py must not import the oracle for its inputs); re-exported
# here because the tests reach it through this module.
# ----------------------------------------------------------------------------------------------
from clairs_to_b200.synth_weights import (PREDICT_CVT, aff_state_dict_shapes, neg_state_dict_shapes,    # noqa: E402,F401
                                          synth_state_dict)
