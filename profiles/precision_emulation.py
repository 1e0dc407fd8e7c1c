"""CPU emulation of the split-precision schemes considered for the tensor-core contractions ("Precision
decision"): the NEG network (BiGRU + heads) with every matmul operand rounded as the hardware would see it, products
accumulated in fp64, against the fp64 forward of the oracle.  Prints max |d logit| per scheme for two weight seeds:
fp32 ~1e-6, tf32x1 ~2e-3, tf32x2 (weights rounded once) ~1e-3, tf32x3 ~1e-6, bf16x3 ~2.5e-5, bf16x6 ~1e-6.
Run here (no GPU needed): python profiles/precision_emulation.py"""
import sys, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from oracle import nn_oracle, pileup_oracle, posterior_oracle
from clairs_to_b200 import synth
import torch.nn.functional as F

def tf32_rn(x):
    xi = x.view(torch.int32)
    r = ((xi + 0x1000) & ~0x1FFF)
    return r.view(torch.float32)
def tf32_tr(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)
def bf(x): return x.to(torch.bfloat16).to(torch.float32)

MODE='fp32'
def mm(a, w):   # a [.., K], w [N, K]
    if MODE=='fp32': return a @ w.t()
    a64=lambda t: t.double()
    if MODE=='tf32x1':
        return (a64(tf32_tr(a)) @ a64(tf32_rn(w)).t()).float()
    if MODE=='tf32x3':
        ah=tf32_tr(a); al=a-ah; wh=tf32_rn(w); wl=tf32_rn(w-wh)
        return (a64(ah)@a64(wh).t() + a64(tf32_tr(al))@a64(wh).t() + a64(ah)@a64(wl).t()).float()
    if MODE=='tf32x2':
        ah=tf32_tr(a); al=a-ah; wh=tf32_rn(w)
        return (a64(ah)@a64(wh).t() + a64(tf32_tr(al))@a64(wh).t()).float()
    if MODE=='bf16x3':
        ah=bf(a); am=bf(a-ah); wh=bf(w); wm=bf(w-wh)
        return (a64(ah)@a64(wh).t() + a64(am)@a64(wh).t() + a64(ah)@a64(wm).t()).float()
    if MODE=='bf16x4':
        ah=bf(a); am=bf(a-ah); wh=bf(w); wm=bf(w-wh)
        return (a64(ah)@a64(wh).t() + a64(am)@a64(wh).t() + a64(ah)@a64(wm).t()+ a64(am)@a64(wm).t()).float()
    if MODE=='bf16x6':
        ah=bf(a); am=bf(a-ah); al=bf(a-ah-am); wh=bf(w); wm=bf(w-wh); wl=bf(w-wh-wm)
        return (a64(ah)@a64(wh).t() + a64(am)@a64(wh).t() + a64(ah)@a64(wm).t()+ a64(am)@a64(wm).t()+a64(al)@a64(wh).t()+a64(ah)@a64(wl).t()).float()
    raise

def gru_dir(x, w_ih, w_hh, b_ih, b_hh, reverse):
    b,t,_=x.shape; H=w_hh.shape[1]
    gi = mm(x, w_ih) + b_ih
    h = x.new_zeros(b,H); out=x.new_empty(b,t,H)
    for s in (range(t-1,-1,-1) if reverse else range(t)):
        gh = mm(h, w_hh)
        r=torch.sigmoid(gi[:,s,:H]+gh[:,:H]+b_hh[:H]); z=torch.sigmoid(gi[:,s,H:2*H]+gh[:,H:2*H]+b_hh[H:2*H])
        n=torch.tanh(gi[:,s,2*H:]+r*(gh[:,2*H:]+b_hh[2*H:]))
        h=(1-z)*n+z*h; out[:,s]=h
    return out
def neg(x, sd):
    y=x
    for name in ('lstm','lstm_2'):
        f=gru_dir(y, sd[name+'.weight_ih_l0'], sd[name+'.weight_hh_l0'], sd[name+'.bias_ih_l0'], sd[name+'.bias_hh_l0'], False)
        bwd=gru_dir(y, sd[name+'.weight_ih_l0_reverse'], sd[name+'.weight_hh_l0_reverse'], sd[name+'.bias_ih_l0_reverse'], sd[name+'.bias_hh_l0_reverse'], True)
        y=torch.cat([f,bwd],-1)
    feat=y.reshape(y.shape[0],-1)
    h=F.selu(mm(feat, sd['fc1.weight'])+sd['fc1.bias'])
    outs=[]
    for n in nn_oracle.head_names(sd, True):
        yy=F.selu(mm(h, sd[n+'_fc2.weight'])+sd[n+'_fc2.bias'])
        outs.append(F.selu(F.linear(yy, sd[n+'_fc3.weight'], sd[n+'_fc3.bias'])))
    return torch.stack(outs,1)

n=256
(aff,aa),(ng,na)=synth.synth_pair(n, 77, 'ont', depth_mean=60, depth_hi=140)
import ctypes
# build input via oracle on CPU (slow python) -- instead use random count-like input
rng=np.random.default_rng(5)
x=np.zeros((n,33,34),np.float32)
depth=rng.integers(10,50,size=(n,33))
for i in range(n):
    for p in range(33):
        d=depth[i,p]; refb=rng.integers(0,4); f=rng.binomial(d,0.5)
        v=np.zeros(34); v[refb]=-f; v[9+refb]=-(d-f)
        alt=rng.integers(0,4); k=rng.binomial(d,0.1); v[alt]+=k
        x[i,p]=v
x=torch.from_numpy(x)
for seed in (204, 206):
    sd=nn_oracle.synth_state_dict(nn_oracle.neg_state_dict_shapes(4), seed)
    ref=nn_oracle.neg_forward(x.numpy(), sd, dtype=torch.float64)
    for m in ('fp32','tf32x1','tf32x2','tf32x3','bf16x3','bf16x4','bf16x6'):
        MODE=m
        with torch.no_grad(): o=neg(x, sd)
        print(seed, m, float((o.double()-ref).abs().max()), float(ref.abs().max()))
