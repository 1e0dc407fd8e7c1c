"""Checkpoint -> flat fp32 weight blobs for the CUDA engine.

The reference stores whole pickled modules, ``{'model_acgt': CvT...}`` / ``{'model_nacgt': BiGRU...}``
(clairs/predict.py:512-568); hyper-parameters live only in tensor shapes, so everything here is
derived from the ``state_dict``.  The segment order below is walked identically
by ``csrc/engine.cu`` (``aff_load`` / ``neg_load``); every segment is padded to 8 floats
(so that the engine's bf16 copies of the same segments stay 16-byte aligned for TMA).

Host-side folding (exact up to fp32 rounding of the folded constants):
  * 3x3 kernels act on H=1 maps -> only the middle kernel row is kept (clairs/model.py:195, 93);
  * eval BatchNorm (clairs/model.py:96): scale folded into the depth-wise taps, shift folded into a
    bias of the following 1x1 conv;
  * attention scale 64^-0.5 = 0.125 (clairs/model.py:108) folded into the q projection (a power of two);
  * GRU b_hh of the r and z gates folded into the input-projection bias (b_hn must stay separate);
  * fc1 columns of AFF permuted from channel-major (c*W + w) to the engine's position-major (w*C + c).
"""

from __future__ import annotations

import numpy as np

DIM_HEAD = 64
BN_EPS = 1e-5
N_POS, N_CH = 33, 34
SEG_ALIGN = 8          # csrc/engine.cuh SEG_ALIGN


def _np(t):
    if hasattr(t, "detach"):
        t = t.detach().cpu().numpy()
    return np.asarray(t, dtype=np.float64)


class _Blob:
    def __init__(self):
        self.parts = []

    def add(self, arr):
        a = np.ascontiguousarray(arr, dtype=np.float32).reshape(-1)
        pad = (-a.size) % SEG_ALIGN
        if pad:
            a = np.concatenate([a, np.zeros(pad, dtype=np.float32)])
        self.parts.append(a)

    def finish(self):
        return np.ascontiguousarray(np.concatenate(self.parts))


def _head_names(sd, negational):
    names = ['a', 'c', 'g', 't', 'i', 'd']
    if negational:
        names = ['n' + n for n in names]
    return [n for n in names if n + '_fc2.weight' in sd]


def _add_heads(blob, sd, names, fc1_w):
    blob.add(fc1_w)
    blob.add(_np(sd['fc1.bias']))
    blob.add(np.concatenate([_np(sd[n + '_fc2.weight']) for n in names], axis=0))
    blob.add(np.concatenate([_np(sd[n + '_fc2.bias']) for n in names], axis=0))
    blob.add(np.stack([_np(sd[n + '_fc3.weight']) for n in names], axis=0))
    blob.add(np.stack([_np(sd[n + '_fc3.bias']) for n in names], axis=0))


def _fold_dw_bn_pw(sd, prefix, scale_out=1.0):
    """DepthWiseConv2d (clairs/model.py:91-100) -> (taps [3, C], pw [O, C], bias [O])."""
    dw = _np(sd[prefix + '.net.0.weight'])[:, 0, 1, :]                    # [C, 3] middle row
    g, b = _np(sd[prefix + '.net.1.weight']), _np(sd[prefix + '.net.1.bias'])
    mu, var = _np(sd[prefix + '.net.1.running_mean']), _np(sd[prefix + '.net.1.running_var'])
    s = g / np.sqrt(var + BN_EPS)
    shift = b - mu * s
    taps = (dw * s[:, None]).T                                             # [3, C]
    pw = _np(sd[prefix + '.net.2.weight'])[:, :, 0, 0] * scale_out         # [O, C]
    bias = pw @ shift
    return taps, pw, bias


def export_aff(sd):
    """state_dict of CvT / CvT_Indel -> (blob float32[n], cfg int32[2 + 3*stages])."""
    names = _head_names(sd, negational=False)
    stages = [s for s in (1, 2, 3) if 'layer%d.0.weight' % s in sd]
    cfg = [len(names), len(stages)]
    blob = _Blob()
    width, c_last = N_POS, N_CH
    for s in stages:
        L = 'layer%d' % s
        w = _np(sd[L + '.0.weight'])                                       # [C, Cin, 3, 3]
        c, cin = w.shape[0], w.shape[1]
        depth = 0
        while '%s.2.layers.%d.0.norm.g' % (L, depth) in sd:
            depth += 1
        inner = sd['%s.2.layers.0.0.fn.to_q.net.2.weight' % L].shape[0] if depth else DIM_HEAD
        cfg += [c, inner // DIM_HEAD, depth]
        blob.add(w[:, :, 1, :].transpose(0, 2, 1).reshape(c, 3 * cin))     # k = tap*Cin + ci
        blob.add(_np(sd[L + '.0.bias']))
        blob.add(_np(sd[L + '.1.g']))
        blob.add(_np(sd[L + '.1.b']))
        for d in range(depth):
            p = '%s.2.layers.%d' % (L, d)
            blob.add(_np(sd[p + '.0.norm.g']))
            blob.add(_np(sd[p + '.0.norm.b']))
            taps, pw, bias = _fold_dw_bn_pw(sd, p + '.0.fn.to_q', scale_out=DIM_HEAD ** -0.5)
            blob.add(taps); blob.add(pw); blob.add(bias)
            taps, pw, bias = _fold_dw_bn_pw(sd, p + '.0.fn.to_kv')
            blob.add(taps); blob.add(pw); blob.add(bias)
            blob.add(_np(sd[p + '.0.fn.to_out.0.weight'])[:, :, 0, 0])
            blob.add(_np(sd[p + '.0.fn.to_out.0.bias']))
            blob.add(_np(sd[p + '.1.norm.g']))
            blob.add(_np(sd[p + '.1.norm.b']))
            blob.add(_np(sd[p + '.1.fn.net.0.weight'])[:, :, 0, 0])
            blob.add(_np(sd[p + '.1.fn.net.0.bias']))
            blob.add(_np(sd[p + '.1.fn.net.3.weight'])[:, :, 0, 0])
            blob.add(_np(sd[p + '.1.fn.net.3.bias']))
        width = (width + 1) // 2
        c_last = c
    fc1 = _np(sd['fc1.weight'])                                            # [128, C*W], index c*W + w
    assert fc1.shape[1] == c_last * width, (fc1.shape, c_last, width)
    fc1 = fc1.reshape(fc1.shape[0], c_last, width).transpose(0, 2, 1).reshape(fc1.shape[0], -1)
    _add_heads(blob, sd, names, fc1)
    return blob.finish(), np.asarray(cfg, dtype=np.int32)


def export_neg(sd):
    """state_dict of BiGRU_NACGT / _Indel -> (blob float32[n], cfg int32[4])."""
    names = _head_names(sd, negational=True)
    blob = _Blob()
    hiddens = []
    for name in ('lstm', 'lstm_2'):
        wih, bih, whh_t, bhn = [], [], [], []
        for suf in ('', '_reverse'):
            w_ih = _np(sd['%s.weight_ih_l0%s' % (name, suf)])              # [3H, in]  gates r|z|n
            w_hh = _np(sd['%s.weight_hh_l0%s' % (name, suf)])              # [3H, H]
            b_ih = _np(sd['%s.bias_ih_l0%s' % (name, suf)])
            b_hh = _np(sd['%s.bias_hh_l0%s' % (name, suf)])
            h = w_hh.shape[1]
            b = b_ih.copy()
            b[:2 * h] += b_hh[:2 * h]
            wih.append(w_ih); bih.append(b); whh_t.append(w_hh.T.copy()); bhn.append(b_hh[2 * h:])
        hiddens.append(h)
        blob.add(np.concatenate(wih, axis=0))
        blob.add(np.concatenate(bih, axis=0))
        blob.add(np.stack(whh_t, axis=0))
        blob.add(np.stack(bhn, axis=0))
    in_dim = sd['lstm.weight_ih_l0'].shape[1]
    _add_heads(blob, sd, names, _np(sd['fc1.weight']))
    cfg = [len(names), int(in_dim), hiddens[0], hiddens[1]]
    return blob.finish(), np.asarray(cfg, dtype=np.int32)


class _ModuleShell:
    """Stand-in for a reference class (``clairs.model.CvT`` ...) when the reference package is not importable: pickle
    only restores attributes (``__init__`` is never called), and all that is needed from the restored object graph are
    the ``_parameters`` / ``_buffers`` / ``_modules`` dictionaries torch.nn.Module keeps."""


def _shell_class(module, name):
    import torch
    return type(name, (torch.nn.Module,), {"__module__": module, "__doc__": _ModuleShell.__doc__})


def _tolerant_pickle_module():
    """A ``pickle_module`` for torch.load whose Unpickler resolves classes of missing ``clairs.*`` / ``shared.*``
    modules to empty torch.nn.Module subclasses."""
    import pickle
    import types

    class Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            try:
                return super().find_class(module, name)
            except (ImportError, AttributeError):
                if module.split('.')[0] in ('clairs', 'shared', 'src'):
                    return _shell_class(module, name)
                raise

    shim = types.ModuleType("cto_tolerant_pickle")
    shim.Unpickler = Unpickler
    shim.load = lambda f, **kw: Unpickler(f, **kw).load()
    for k in ("__name__", "dumps", "dump", "loads", "PickleError", "UnpicklingError", "PicklingError", "HIGHEST_PROTOCOL", "DEFAULT_PROTOCOL"):
        if hasattr(pickle, k) and k != "__name__":
            setattr(shim, k, getattr(pickle, k))
    return shim


def state_dict_from_checkpoint(path, key):
    """Load a reference checkpoint: ``torch.load(..., weights_only=False)[key]`` is a whole pickled ``clairs.model``
    module (clairs/predict.py:513-517); a bare ``state_dict`` under the key is accepted too.  Inside a ClairS-TO checkout
    ``clairs.model`` is importable and the pickle loads as it does there; elsewhere the classes are resolved to empty
    torch.nn.Module shells, which is enough to read the parameters (hyper-parameters are derived from tensor shapes)."""
    import torch
    try:
        obj = torch.load(path, map_location='cpu', weights_only=False)
    except (ImportError, AttributeError, ModuleNotFoundError):
        obj = torch.load(path, map_location='cpu', weights_only=False, pickle_module=_tolerant_pickle_module())
    model = obj[key] if isinstance(obj, dict) and key in obj else obj
    return model.state_dict() if hasattr(model, 'state_dict') else model


def likelihood_tables(path_or_array, n_heads):
    """likelihood_matrix.txt (clairs/call_variants.py:655-796) -> double[n_heads, 122]:
    100 matrix entries, 11 AFF edges [0, e0..e8, 1], 11 NEG edges."""
    data = np.loadtxt(path_or_array) if isinstance(path_or_array, str) else np.asarray(path_or_array, dtype=np.float64)
    need = 12 * n_heads
    if data.ndim != 2 or data.shape[0] < need or data.shape[1] != 10:
        raise ValueError("likelihood matrix: expected >= %d rows of 10 columns, got %r" % (need, data.shape))
    out = np.zeros((n_heads, 122), dtype=np.float64)
    for h in range(n_heads):
        out[h, :100] = data[10 * h:10 * h + 10].reshape(-1)
        for k, row in enumerate((10 * n_heads + 2 * h, 10 * n_heads + 2 * h + 1)):
            edges = np.concatenate([[0.0], data[row, :-1], [1.0]])
            out[h, 100 + 11 * k:111 + 11 * k] = edges
    return out
