"""Seeded random-init weights in the reference's ``state_dict`` key layout.

Trained checkpoints are not available offline, so the benchmark, ``smoke()`` and the tests run the engine (and the
oracle) on deterministic synthetic weights of the reference architectures: CvT / CvT_Indel with the
clairs/predict.py:520-553 hyper-parameters and BiGRU_NACGT / _Indel (clairs/model.py:387-467).
"""

from __future__ import annotations

import math

import torch

DIM_HEAD = 64          # clairs/model.py:103

PREDICT_CVT = dict(s1=(16, 1, 1), s2=(64, 3, 2), s3=(128, 4, 3))   # (emb_dim, heads, depth), predict.py:520-553


def aff_state_dict_shapes(n_heads=4, cfg=None):
    cfg = cfg or PREDICT_CVT
    shapes = {}
    cin = 34
    for si, key in enumerate(('s1', 's2', 's3'), start=1):
        c, heads, depth = cfg[key]
        inner = heads * DIM_HEAD
        L = 'layer%d' % si
        shapes[L + '.0.weight'] = (c, cin, 3, 3)
        shapes[L + '.0.bias'] = (c,)
        shapes[L + '.1.g'] = (1, c, 1, 1)
        shapes[L + '.1.b'] = (1, c, 1, 1)
        for d in range(depth):
            p = '%s.2.layers.%d' % (L, d)
            shapes[p + '.0.norm.g'] = (1, c, 1, 1)
            shapes[p + '.0.norm.b'] = (1, c, 1, 1)
            for proj, mult in (('to_q', 1), ('to_kv', 2)):
                q = '%s.0.fn.%s.net' % (p, proj)
                shapes[q + '.0.weight'] = (c, 1, 3, 3)
                shapes[q + '.1.weight'] = (c,)
                shapes[q + '.1.bias'] = (c,)
                shapes[q + '.1.running_mean'] = (c,)
                shapes[q + '.1.running_var'] = (c,)
                shapes[q + '.2.weight'] = (inner * mult, c, 1, 1)
            shapes[p + '.0.fn.to_out.0.weight'] = (c, inner, 1, 1)
            shapes[p + '.0.fn.to_out.0.bias'] = (c,)
            shapes[p + '.1.norm.g'] = (1, c, 1, 1)
            shapes[p + '.1.norm.b'] = (1, c, 1, 1)
            shapes[p + '.1.fn.net.0.weight'] = (4 * c, c, 1, 1)
            shapes[p + '.1.fn.net.0.bias'] = (4 * c,)
            shapes[p + '.1.fn.net.3.weight'] = (c, 4 * c, 1, 1)
            shapes[p + '.1.fn.net.3.bias'] = (c,)
        cin = c
    width = 33
    for _ in range(3):
        width = math.ceil(width / 2)
    shapes['fc1.weight'] = (128, cin * width)
    shapes['fc1.bias'] = (128,)
    shapes['fc2.weight'] = (2, 128)            # unused by forward (M:215)
    shapes['fc2.bias'] = (2,)
    for n in ['a', 'c', 'g', 't', 'i', 'd'][:n_heads]:
        shapes[n + '_fc2.weight'] = (128, 128)
        shapes[n + '_fc2.bias'] = (128,)
        shapes[n + '_fc3.weight'] = (2, 128)
        shapes[n + '_fc3.bias'] = (2,)
    return shapes


def neg_state_dict_shapes(n_heads=4):
    shapes = {}
    for name, cin, hid in (('lstm', 34, 128), ('lstm_2', 256, 192)):
        for suf in ('', '_reverse'):
            shapes['%s.weight_ih_l0%s' % (name, suf)] = (3 * hid, cin)
            shapes['%s.weight_hh_l0%s' % (name, suf)] = (3 * hid, hid)
            shapes['%s.bias_ih_l0%s' % (name, suf)] = (3 * hid,)
            shapes['%s.bias_hh_l0%s' % (name, suf)] = (3 * hid,)
    shapes['fc1.weight'] = (128, 33 * 384)
    shapes['fc1.bias'] = (128,)
    shapes['fc2.weight'] = (128, 128)          # unused by forward (M:424)
    shapes['fc2.bias'] = (128,)
    for n in ['na', 'nc', 'ng', 'nt', 'ni', 'nd'][:n_heads]:
        shapes[n + '_fc2.weight'] = (128, 128)
        shapes[n + '_fc2.bias'] = (128,)
        shapes[n + '_fc3.weight'] = (2, 128)
        shapes[n + '_fc3.bias'] = (2,)
    return shapes


def synth_state_dict(shapes, seed, gain=1.0):
    """Seeded weights: fan-in scaled normals; LayerNorm g/b and BatchNorm statistics are
    randomised too (defaults 1/0/0/1 would hide folding bugs)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    sd = {}
    for key, shape in shapes.items():
        if key.endswith('running_var'):
            v = rng.uniform(0.5, 2.0, size=shape)
        elif key.endswith('running_mean'):
            v = rng.normal(0.0, 0.3, size=shape)
        elif key.endswith('.g') or key.endswith('net.1.weight'):
            v = rng.uniform(0.6, 1.4, size=shape)
        elif key.endswith('.b') or key.endswith('bias'):
            v = rng.normal(0.0, 0.1, size=shape)
        else:
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else int(shape[0])
            if len(shape) == 4 and shape[2] == 3:
                fan_in = shape[1] * 3                      # only the middle kernel row is live
            v = rng.normal(0.0, gain / math.sqrt(max(fan_in, 1)), size=shape)
        sd[key] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
    return sd
