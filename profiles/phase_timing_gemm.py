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
for (m,n,k) in [(161024,64,16),(161024,64,64),(9472*33,1152,256),(9472*5,512,128)]:
    a = torch.randn(m,k,device='cuda'); w = torch.randn(n,k,device='cuda')/k**0.5; b = torch.randn(n,device='cuda')
    for dbg in (0, 1):
        lib.cto_debug_set(dbg)
        for _ in range(3): gemm_nt(a,w,b,None,0,True)
        torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): gemm_nt(a,w,b,None,0,True)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1)*100
        if dbg == 0: print("shape",m,n,k,"time/launch %.1f us (incl. weight split)"%us)
    t = buf.cpu().tolist()
    tiles_per_cta = t[20]/148.0
    print("  tiles/cta %.1f kb %d ; cycles per tile (CTA0):"%(tiles_per_cta, t[21]))
    for i,nm in enumerate(names):
        if nm != "-": print("    %-20s %8.0f"%(nm, t[i]/max(tiles_per_cta,1)))
