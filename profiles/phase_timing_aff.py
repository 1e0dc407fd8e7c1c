"""Per-phase clock64() counters of CTA 0 of the fused CvT layer kernel (csrc/aff_fused.cu), per stage."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clairs_to_b200 import _lib
from clairs_to_b200.engine import Engine
from oracle import nn_oracle

lib = _lib.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 33333
aff_sd = nn_oracle.synth_state_dict(nn_oracle.aff_state_dict_shapes(4), 104)
neg_sd = nn_oracle.synth_state_dict(nn_oracle.neg_state_dict_shapes(4), 204)
eng = Engine(aff_sd, neg_sd, max_batch=n)
buf = torch.zeros(32, dtype=torch.int64, device='cuda')
MMA = ["issue/other", "wait w_full", "wait pc_ready", "wait ab_ready", "wait qkv_free", "wait y2_ready", "wait hid_free"]
GA = ["x load", "LN1+dwconv", "wait qkv_full", "attention math", "wait pc_free", "wait x_ready", "LN2", "wait hid_full", "gelu+store", "wait layer_done", "x store"]
for stage, W, Cc, depth in ((1, 9, 64, 2), (2, 5, 128, 3)):
    x = torch.randn(n, W, Cc, device='cuda')
    for rep in range(3):
        lib.cto_debug_timing_fused(C.c_void_p(buf.data_ptr()) if rep == 2 else None)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.cto_aff_stage_layers(eng.handle, stage, C.c_void_p(x.data_ptr()), n, None), "layers")
        e1.record(); torch.cuda.synchronize()
    lib.cto_debug_timing_fused(None)
    t = buf.cpu().tolist()
    lt = max(t[7], 1)
    print("stage %d: C=%d W=%d depth=%d, %d candidates: %.3f ms (timed launch); CTA 0 ran %d layer-tiles; cycles per layer-tile:" % (stage, Cc, W, depth, n, e0.elapsed_time(e1), lt))
    for i, nm in enumerate(MMA):
        print("    mma  %-16s %9.0f" % (nm, t[i] / lt))
    for i, nm in enumerate(GA):
        print("    cmpA %-16s %9.0f" % (nm, t[8 + i] / lt))
