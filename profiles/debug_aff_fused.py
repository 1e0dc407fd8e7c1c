"""Bisection harness for the fused CvT layer kernel (csrc/aff_fused.cu): runs the layers of one stage on a random
residual stream through cto_aff_stage_layers in fused mode and in kernel-per-op mode and compares both with the oracle,
with selected weights zeroed so that single blocks (TMEM round trip / feed-forward / attention) can be isolated."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clairs_to_b200 import _lib
from clairs_to_b200.engine import Engine
from oracle import nn_oracle


def oracle_layers(x, sd, name):
    """x [n, W, C] -> layers of stage `name` (no embed conv / LN)."""
    sd = {k: v.double() for k, v in sd.items()}
    t = torch.from_numpy(x).double().permute(0, 2, 1)
    d = 0
    while '%s.2.layers.%d.0.norm.g' % (name, d) in sd:
        p = '%s.2.layers.%d' % (name, d)
        t = nn_oracle._attention(nn_oracle._channel_ln(t, sd[p + '.0.norm.g'], sd[p + '.0.norm.b']), sd, p + '.0.fn') + t
        t = nn_oracle._feed_forward(nn_oracle._channel_ln(t, sd[p + '.1.norm.g'], sd[p + '.1.norm.b']), sd, p + '.1.fn') + t
        d += 1
    return t.permute(0, 2, 1).numpy()


def run(eng, stage, x, mode):
    eng.set_tensor_cores(mode)
    t = torch.from_numpy(x).cuda().contiguous()
    _lib.check(eng.lib.cto_aff_stage_layers(eng.handle, stage, C.c_void_p(t.data_ptr()), x.shape[0], None), "stage_layers")
    torch.cuda.synchronize()
    st = eng.fused_status()
    return t.cpu().numpy(), st


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    variants = sys.argv[2].split(',') if len(sys.argv) > 2 else ['ident', 'ff', 'att', 'full']
    neg_sd = nn_oracle.synth_state_dict(nn_oracle.neg_state_dict_shapes(4), 204)
    for stage, name, W, Cc in ((1, 'layer2', 9, 64), (2, 'layer3', 5, 128)):
        for variant in variants:
            sd = nn_oracle.synth_state_dict(nn_oracle.aff_state_dict_shapes(4), 104)
            for k in list(sd):
                if not k.startswith(name + '.2.layers'):
                    continue
                if variant in ('ident', 'ff') and '.0.fn.to_out.0.' in k:
                    sd[k] = torch.zeros_like(sd[k])
                if variant in ('ident', 'att') and '.1.fn.net.3.' in k:
                    sd[k] = torch.zeros_like(sd[k])
            eng = Engine(sd, neg_sd, max_batch=max(n, 64))
            x = np.random.default_rng(stage).normal(0, 1.5, size=(n, W, Cc)).astype(np.float32)
            want = oracle_layers(x, sd, name)
            for mode in (2, 1):
                got, st = run(eng, stage, x, mode)
                err = np.abs(got - want)
                print("stage %d (%s) variant %-5s mode %d: max err %.3e  mean %.3e  status %s" %
                      (stage, name, variant, mode, err.max(), err.mean(), st[:4].tolist()), flush=True)
                if mode == 1 and err.max() > 1e-3:
                    bad = np.argwhere(err > 1e-3)
                    cands = np.unique(bad[:, 0]); ws = np.unique(bad[:, 1]); cs = np.unique(bad[:, 2])
                    print("   bad: %d elements; candidates %s; positions %s; channels %d distinct (first %s)" %
                          (len(bad), cands[:20].tolist(), ws.tolist(), len(cs), cs[:16].tolist()))
                    r = bad[0]
                    print("   e.g. [%d,%d,%d] got %.5f want %.5f in %.5f" % (r[0], r[1], r[2], got[tuple(r)], want[tuple(r)], x[tuple(r)]))
            eng.close()


if __name__ == '__main__':
    main()
