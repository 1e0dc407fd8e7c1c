"""Shared-memory-bandwidth model of the bf16x3 GEMM kernel: per 128 x BN tile, bytes through shared
memory (MMA operand fetch, TMA writes of A / W, optional converter pass, epilogue staging write + TMA-store read) divided by
128 B/clk, against the pure MMA time; summed over the GEMMs of both networks for 100 000 candidates in 9 472-candidate
chunks.  It reproduced the measured proj2 / fc1 times within 15 % and showed that the AFF 1x1 GEMMs were 2-3x above their
bound (launch + converter), which motivated the split-plane operands.  Run: python profiles/gemm_smem_model.py"""
# smem-bandwidth model of the GEMM kernel: bytes through shared memory per tile / 128 B/clk
import math
N=100000
def gemms():
    L=[]
    # AFF stages: (C, heads, depth, W, Wkv)
    for (c,h,d,w,wkv) in ((16,1,1,17,9),(64,3,2,9,5),(128,4,3,5,3)):
        inner=64*h
        for _ in range(d):
            L+=[('q',N*w,inner,c),('kv',N*wkv,2*inner,c),('out',N*w,c,inner),('ff1',N*w,4*c,c),('ff2',N*w,c,4*c)]
    L+=[('afc1',N,128,640),('afc2',N,512,128)]
    L+=[('proj1',N*33,768,40),('proj2',N*33,1152,256),('nfc1',N,128,12672),('nfc2',N,512,128)]
    return L
def tile_time(m,n,k,mode):
    if n<64: return None
    chunk=9472*(m//N)
    # per chunk launch
    launches=N/9472
    bn = 128 if (n%128==0 and math.ceil(chunk/128)*(n//128)>=148) else 64
    pair=False
    if mode in('bn256','pair256') and n%256==0: bn=256
    tiles=math.ceil(chunk/128)*(n//bn)
    ksteps=math.ceil(k/16)
    kb=math.ceil(k/64)
    presplit = mode in ('presplit','bn256','pair256')
    if mode=='pair256' and bn==256:
        mma_bytes=ksteps*(3*(4096+4096))   # A 128x16x2 + Whalf 128x16x2
        w_tma=kb*2*128*128
    else:
        mma_bytes=ksteps*(3*(4096+bn*32))
        w_tma=kb*2*bn*128
    a_tma=kb*32768 if not presplit else kb*32768
    conv=0 if presplit else kb*65536
    epi=128*bn*4*2
    tot=mma_bytes+w_tma+a_tma+conv+epi
    mma_clk=ksteps*3*(128*bn*16/4096)
    clk=max(tot/128, mma_clk)
    waves=math.ceil(tiles/148)
    t=waves*clk/1.9e9*launches
    return t*1e3, bn, tot/128, mma_clk
for mode in ('now','presplit','bn256','pair256'):
    tot=0; aff=0
    for (name,m,n,k) in gemms():
        r=tile_time(m,n,k,mode)
        if r is None: continue
        tot+=r[0]
        if name in('q','kv','out','ff1','ff2'): aff+=r[0]
        if mode=='now' or name in ('proj2',): print(mode,name,m,n,k,'%.2f ms bn=%d smemclk=%d mmaclk=%d'%r)
    print(mode,'TOTAL %.2f ms  (aff 1x1: %.2f)'%(tot,aff))
