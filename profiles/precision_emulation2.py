"""CPU emulation (round 2) of two-product split schemes for the tensor-core contractions of BOTH networks.

Question: can the 3-product bf16x3 scheme (act hi*W hi + act mid*W hi + act hi*W mid) drop
to two products by keeping the WEIGHT in one 11-bit-significand number (fp16, same significand as TF32) while the
activation stays an exact-to-16-bit bf16 hi + mid pair?   kind::f16 MMAs take an f16 or a bf16 operand on either
side, so  (a_hi + a_mid) * fp16(w)  is two MMAs per k-step and half the weight bytes in shared memory.

Every matmul operand is rounded as the hardware would see it, products accumulate in fp64, the comparison is the
fp64 forward of the oracle.  Elementwise math stays fp32 (as in the kernels).
Run here (no GPU needed):  python profiles/precision_emulation2.py [n_candidates]
"""
import math
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nn_oracle  # noqa: E402


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def f16(x):
    return x.to(torch.float16).to(torch.float32)


def tf32_rn(x):
    return ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


MODE = 'fp32'


def mm(a, w):
    """a [.., K] (activation), w [N, K] (weight) -> a @ w^T under the current scheme."""
    d = lambda t: t.double()
    if MODE == 'fp32':
        return a @ w.t()
    if MODE == 'bf16x3':
        ah = bf(a); am = bf(a - ah); wh = bf(w); wm = bf(w - wh)
        return (d(ah) @ d(wh).t() + d(am) @ d(wh).t() + d(ah) @ d(wm).t()).float()
    if MODE == 'b2_wf16':          # act bf16 hi+mid, weight fp16: 2 products
        ah = bf(a); am = bf(a - ah); wh = f16(w)
        return (d(ah) @ d(wh).t() + d(am) @ d(wh).t()).float()
    if MODE == 'b2_wbf16':         # act bf16 hi+mid, weight bf16: 2 products
        ah = bf(a); am = bf(a - ah); wh = bf(w)
        return (d(ah) @ d(wh).t() + d(am) @ d(wh).t()).float()
    if MODE == 'f2_wf16':          # act fp16 hi+mid, weight fp16
        ah = f16(a); am = f16(a - ah); wh = f16(w)
        return (d(ah) @ d(wh).t() + d(am) @ d(wh).t()).float()
    if MODE == 'b1f1':             # act bf16 hi only... (reference point: one product)
        return (d(bf(a)) @ d(f16(w)).t()).float()
    if MODE == 'wtf32':            # exact activations, TF32-rounded weights (the SURVEY 0.7 probe)
        return (d(a) @ d(tf32_rn(w)).t()).float()
    raise ValueError(MODE)


# ---- NEG ---------------------------------------------------------------------------------------
def gru_dir(x, w_ih, w_hh, b_ih, b_hh, reverse):
    b, t, _ = x.shape
    H = w_hh.shape[1]
    gi = mm(x, w_ih) + b_ih
    h = x.new_zeros(b, H)
    out = x.new_empty(b, t, H)
    for s in (range(t - 1, -1, -1) if reverse else range(t)):
        gh = mm(h, w_hh)
        r = torch.sigmoid(gi[:, s, :H] + gh[:, :H] + b_hh[:H])
        z = torch.sigmoid(gi[:, s, H:2 * H] + gh[:, H:2 * H] + b_hh[H:2 * H])
        n = torch.tanh(gi[:, s, 2 * H:] + r * (gh[:, 2 * H:] + b_hh[2 * H:]))
        h = (1 - z) * n + z * h
        out[:, s] = h
    return out


def heads(feat, sd, negational):
    h = F.selu(mm(feat, sd['fc1.weight']) + sd['fc1.bias'])
    outs = []
    for n in nn_oracle.head_names(sd, negational):
        y = F.selu(mm(h, sd[n + '_fc2.weight']) + sd[n + '_fc2.bias'])
        outs.append(F.selu(F.linear(y, sd[n + '_fc3.weight'], sd[n + '_fc3.bias'])))
    return torch.stack(outs, 1)


def neg(x, sd):
    y = x
    for name in ('lstm', 'lstm_2'):
        f = gru_dir(y, sd[name + '.weight_ih_l0'], sd[name + '.weight_hh_l0'], sd[name + '.bias_ih_l0'], sd[name + '.bias_hh_l0'], False)
        b = gru_dir(y, sd[name + '.weight_ih_l0_reverse'], sd[name + '.weight_hh_l0_reverse'], sd[name + '.bias_ih_l0_reverse'],
                    sd[name + '.bias_hh_l0_reverse'], True)
        y = torch.cat([f, b], -1)
    return heads(y.reshape(y.shape[0], -1), sd, True)


# ---- AFF ---------------------------------------------------------------------------------------
def ln(x, g, b):        # x [B, W, C]
    mean = x.mean(-1, keepdim=True)
    std = x.var(-1, unbiased=False, keepdim=True).sqrt()
    return (x - mean) / (std + 1e-5) * g.reshape(1, 1, -1) + b.reshape(1, 1, -1)


def conv3(x, w, stride):   # x [B, W, Cin] -> rows of the 3-tap conv, pad 1: [B, Wout, 3*Cin]
    B, W, C = x.shape
    xp = F.pad(x, (0, 0, 1, 1))
    wout = (W + 2 - 3) // stride + 1
    idx = torch.arange(wout) * stride
    return torch.cat([xp[:, idx + t] for t in range(3)], dim=-1)


def dw_bn_pw(y, sd, p, stride):
    dw = sd[p + '.net.0.weight'][:, 0, 1, :]                               # [C, 3]
    s = sd[p + '.net.1.weight'] / torch.sqrt(sd[p + '.net.1.running_var'] + 1e-5)
    shift = sd[p + '.net.1.bias'] - sd[p + '.net.1.running_mean'] * s
    B, W, C = y.shape
    cols = conv3(y, None, stride).reshape(B, -1, 3, C)
    t = (cols * dw.t().reshape(1, 1, 3, C)).sum(2) * s + shift
    return mm(t, sd[p + '.net.2.weight'][:, :, 0, 0])


def aff(x, sd):
    for name in ('layer1', 'layer2', 'layer3'):
        w = sd[name + '.0.weight'][:, :, 1, :]                             # [C, Cin, 3]
        c = w.shape[0]
        x = mm(conv3(x, None, 2), w.permute(0, 2, 1).reshape(c, -1)) + sd[name + '.0.bias']
        x = ln(x, sd[name + '.1.g'], sd[name + '.1.b'])
        d = 0
        while '%s.2.layers.%d.0.norm.g' % (name, d) in sd:
            p = '%s.2.layers.%d' % (name, d)
            y = ln(x, sd[p + '.0.norm.g'], sd[p + '.0.norm.b'])
            q = dw_bn_pw(y, sd, p + '.0.fn.to_q', 1)
            kv = dw_bn_pw(y, sd, p + '.0.fn.to_kv', 2)
            inner = q.shape[-1]
            hds = inner // 64
            B = x.shape[0]
            sp = lambda t: t.reshape(B, t.shape[1], hds, 64).permute(0, 2, 1, 3)
            qh, kh, vh = sp(q), sp(kv[..., :inner]), sp(kv[..., inner:])
            att = torch.softmax(qh @ kh.transpose(-1, -2) * 0.125, -1) @ vh   # CUDA cores, fp32
            att = att.permute(0, 2, 1, 3).reshape(B, -1, inner)
            x = x + mm(att, sd[p + '.0.fn.to_out.0.weight'][:, :, 0, 0]) + sd[p + '.0.fn.to_out.0.bias']
            y = ln(x, sd[p + '.1.norm.g'], sd[p + '.1.norm.b'])
            h = F.gelu(mm(y, sd[p + '.1.fn.net.0.weight'][:, :, 0, 0]) + sd[p + '.1.fn.net.0.bias'])
            x = x + mm(h, sd[p + '.1.fn.net.3.weight'][:, :, 0, 0]) + sd[p + '.1.fn.net.3.bias']
            d += 1
    feat = x.permute(0, 2, 1).reshape(x.shape[0], -1)                      # channel-major flatten
    return heads(feat, sd, False)


def count_like_input(n, seed):
    rng = np.random.default_rng(seed)
    x = np.zeros((n, 33, 34), np.float32)
    depth = rng.integers(10, 50, size=(n, 33))
    for i in range(n):
        for p in range(33):
            d = depth[i, p]; refb = rng.integers(0, 4); f = rng.binomial(d, 0.5)
            v = np.zeros(34); v[refb] = -f; v[9 + refb] = -(d - f)
            alt = rng.integers(0, 4); k = rng.binomial(d, 0.1); v[alt] += k
            x[i, p] = v
    return torch.from_numpy(x)


if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
    x = count_like_input(n, 5)
    modes = ('fp32', 'bf16x3', 'b2_wf16', 'f2_wf16', 'b2_wbf16', 'wtf32', 'b1f1')
    for gain in (0.5, 1.0, 1.5):
        for seed in (4, 6):
            nsd = nn_oracle.synth_state_dict(nn_oracle.neg_state_dict_shapes(4), 200 + seed, gain)
            asd = nn_oracle.synth_state_dict(nn_oracle.aff_state_dict_shapes(4), 100 + seed, gain)
            nref = nn_oracle.neg_forward(x.numpy(), nsd, dtype=torch.float64)
            aref = nn_oracle.aff_forward(x.numpy(), asd, dtype=torch.float64)
            for m in modes:
                MODE = m
                with torch.no_grad():
                    en = float((neg(x, nsd).double() - nref).abs().max())
                    ea = float((aff(x, asd).double() - aref).abs().max())
                print("gain %.1f seed %d %-9s  NEG %.2e (max|logit| %.1f)   AFF %.2e (max|logit| %.1f)"
                      % (gain, seed, m, en, float(nref.abs().max()), ea, float(aref.abs().max())), flush=True)
