import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from clairs_to_b200 import _lib
from clairs_to_b200.engine import Engine
from oracle import nn_oracle
"""Per-phase clock64() counters of the CTA-pair GRU kernel (cluster 0, forward direction, gate warp 2 and the MMA warp).
The last kernel launched before the read is GRU-2 (H=192) of the last chunk."""
lib = _lib.lib()
lib.cto_debug_timing.argtypes = [C.c_void_p]
buf = torch.zeros(128, dtype=torch.int64, device='cuda')      # [0,32) GEMM kernels, [32,64) GRU kernel
aff_sd = nn_oracle.synth_state_dict(nn_oracle.aff_state_dict_shapes(4), 104)
neg_sd = nn_oracle.synth_state_dict(nn_oracle.neg_state_dict_shapes(4), 204)
n = 37888
eng = Engine(aff_sd, neg_sd, max_batch=n)
x = torch.from_numpy(np.random.default_rng(0).integers(-50, 51, size=(n, 33, 34)).astype(np.float32)).cuda()
for dbg in (0, 2, 4, 8, 14, 1):      # 2: no xproj loads, 4: no output stores, 8: no MUFU, 14: all three (timing experiments, wrong results)
    lib.cto_debug_timing(C.c_void_p(buf.data_ptr())); lib.cto_debug_set(dbg)
    for _ in range(2): eng.forward_neg(x)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); eng.forward_neg(x); e1.record(); torch.cuda.synchronize()
    print("dbg", dbg, "neg forward %.3f ms for %d candidates" % (e0.elapsed_time(e1), n))
t = buf.cpu().tolist()[32:]
names = {0: "mma.wait_h_ready", 1: "mma.wait_w_full", 2: "mma.issue+commit", 8: "gate.wait_acc(first pair)", 9: "gate.wait_acc(middle)",
         10: "gate.wait_acc(last pair)", 11: "gate.tmem_ld", 12: "gate.math+stores", 13: "gate.wait_x_stage"}
print("cycles per step (33 steps), H=192:")
for i, nm in names.items(): print("    %-28s %9.0f" % (nm, t[i] / 33.0))
