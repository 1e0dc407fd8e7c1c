import sys, ctypes as C, torch, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from clairs_to_b200 import _lib
from clairs_to_b200.engine import Engine
from oracle import nn_oracle
lib = _lib.lib()
buf = torch.zeros(32, dtype=torch.int64, device='cuda')
eng = Engine(nn_oracle.synth_state_dict(nn_oracle.aff_state_dict_shapes(4), 104), nn_oracle.synth_state_dict(nn_oracle.neg_state_dict_shapes(4), 204), max_batch=9472)
x = torch.randn(9472, 33, 34, device='cuda')
for c, dbg in ((1,0),(2,0)):
    lib.cto_debug_gru_cluster(c); lib.cto_debug_set(dbg)
    eng.forward_neg(x); torch.cuda.synchronize()
    lib.cto_debug_timing(C.c_void_p(buf.data_ptr()))
    eng.forward_neg(x); torch.cuda.synchronize()
    lib.cto_debug_timing(None)
    t = buf[24:30].cpu().tolist()
    names = ["mma.wait_h_ready", "mma.wait_full(W)", "mma.issue+commit", "epi.wait_acc_full", "epi.gate_math", "epi.h_writeback+arrive"]
    print("dbg", dbg, "cluster", c, "(last GRU launch = layer 2, H=192) cycles per step:")
    for n, v in zip(names, t): print("   %-20s %9.0f" % (n, v / 33))
