"""Phase counters of the GEMM kernel on the transposed layer-2 projection (A = W_ih [1152, 256] pre-split, W = the layer-1
output of one engine chunk [33 * 37888, 256], bias per row, n-major tiles), 128 x 128 tiles (mode 11) against 128 x 256 (mode 27) and the CTA-pair kernel with A resident (mode 43, gemm_pair.cu)."""
import sys, ctypes as C, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from clairs_to_b200 import _lib
from clairs_to_b200.engine import gemm_nt
lib = _lib.lib()
lib.cto_debug_timing.argtypes = [C.c_void_p]
buf = torch.zeros(128, dtype=torch.int64, device='cuda')
lib.cto_debug_timing(C.c_void_p(buf.data_ptr()))
names = ["prod.wait_empty","prod.issue","-","-","mma.wait_acc_empty","mma.wait_full","mma.wait_conv","mma.issue+commit",
         "conv.wait_full","conv.math","conv.fence","-","epi.wait_acc_full","epi.tmem_ld","epi.waitgrp+bar","epi.math+sts","epi.fence","epi.bar2"]
m, n, k = 1152, 33 * 37888, 256
a = torch.randn(m, k, device='cuda'); w = torch.randn(n, k, device='cuda') / k ** 0.5; b = torch.randn(m, device='cuda')
for mode in (11, 27, 43):
    for dbg in (0, 1):
        lib.cto_debug_set(dbg)
        for _ in range(2): gemm_nt(a, w, b, None, 0, mode)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): gemm_nt(a, w, b, None, 0, mode)
        e1.record(); torch.cuda.synchronize()
        if dbg == 0: print("mode", mode, "shape", m, n, k, "time/call %.3f ms (incl. operand split passes of the hook)" % (e0.elapsed_time(e1) / 3))
    t = buf.cpu().tolist()
    tiles_per_cta = t[20] / 148.0
    print("  tiles/cta %.1f kb %d ; cycles per tile (CTA0):" % (tiles_per_cta, t[21]))
    for i, nm in enumerate(names):
        if nm != "-": print("    %-20s %8.0f" % (nm, t[i] / max(tiles_per_cta, 1)))
