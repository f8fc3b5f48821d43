#!/usr/bin/env python
"""CPU study behind orcai_calibrate (test infrastructure; uses the oracle).

fp16 rounding of a layer's weights is the same at every pixel, so - unlike activation rounding, which averages out - it
shifts the layer's output coherently by  sum_k dW[k][n] * A[p, k].  Replacing A by its channel mean turns that into a
constant per output channel that can be folded into the (exact, split-fp16) bias.  This script emulates the fast path's
rounding points in the oracle graph and shows the effect of the correction when the means come from the same input,
from uniform-random input and from another recording:

    fp16 (current)                                 max 4.72e-03 mean 1.84e-04
    bias-corrected, mu from the same input         max 1.53e-03 mean 1.28e-04
    bias-corrected, mu from uniform-random input   max 4.04e-03 mean 2.88e-04
    bias-corrected, mu from another recording      max 1.54e-03 mean 1.29e-04
"""
import sys, numpy as np, torch, torch.nn.functional as F
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tools'))
from oracle import network_oracle as no, postprocess_oracle as po, spectrogram_oracle as so
from orcai_b200 import runtime
from orcai_b200.synth import pcm16_to_float, synth_pcm16
from orcai_b200.weights import synthetic_weights
import precision_study as ps
torch.set_num_threads(8)
P,S=runtime.bundled_parameters()
W=synthetic_weights(P,S,seed=1234)
pcm=synth_pcm16(40.0, seed=20251018)
spec,_,_=so.make_spectrogram(pcm16_to_float(pcm),P["spectrogram"])
x=po.cut_snippets(spec,736)[:8]
xr=np.random.default_rng(5).random((2,736,171),dtype=np.float32)
pcm2=synth_pcm16(20.0, seed=4242, calls_per_minute=10.0)
spec2,_,_=so.make_spectrogram(pcm16_to_float(pcm2),P["spectrogram"])
x2=po.cut_snippets(spec2,736)[:3]
f16=torch.float16
def h(t,on=True): return t.to(f16).to(torch.float32) if on else t
f32=torch.float32; T=lambda a: torch.as_tensor(np.asarray(a),dtype=f32)
def bn_fold(prefix):
    g,b,m,v=(T(W[f"{prefix}/{k}"]).double() for k in ("gamma","beta","moving_mean","moving_variance"))
    s=g/torch.sqrt(v+no.BN_EPS); return s,b-m*s
def folded(sp,bnp):
    dw=T(W[f"{sp}/depthwise_kernel"]).double()[...,0]; pw=T(W[f"{sp}/pointwise_kernel"]).double()[0,0]; b=T(W[f"{sp}/bias"]).double()
    s,t=bn_fold(bnp); pws=(pw*s[None,:]).float()
    return dw.float()[:,:,:,None]*pws[None,None], (b*s+t).float()   # (3,3,C,O), bias
@torch.no_grad()
def fwd(xin, mu=None, collect=None):
    """fp16 emulation; mu: dict layer-> per-input-channel mean used for bias correction; collect: dict to fill with means"""
    def sep(a,sp,bnp,key):
        wt,bias=folded(sp,bnp); wq=h(wt)
        if collect is not None: collect[key]=a.mean(dim=(0,2,3))
        if mu is not None and key in mu:
            dW=(wq-wt).sum(dim=(0,1))        # (C,O) summed over taps
            bias=bias-(mu[key][:,None]*dW).sum(0)
        return F.conv2d(a,wq.permute(3,2,0,1).contiguous(),bias,padding=1)
    xx=torch.as_tensor(xin,dtype=f32)[:,None]
    s,t=bn_fold("bn0"); k0=(T(W["conv0/kernel"]).double()*s[None,None,None,:]).float(); b0=(T(W["conv0/bias"]).double()*s+t).float()
    y=h(torch.relu(F.conv2d(h(xx),h(k0).permute(3,2,0,1).contiguous(),b0,padding=1))); prev=y
    for b in range(1,5):
        p=f"block{b}"
        a=h(torch.relu(sep(torch.relu(prev),f"{p}/sep1",f"{p}/bn1",f"{p}s1")))
        z=h(sep(a,f"{p}/sep2",f"{p}/bn2",f"{p}s2"))
        z=no._maxpool_3x2_s2_same(z)
        rk=T(W[f"{p}/res/kernel"]); rq=h(rk); rb=T(W[f"{p}/res/bias"])
        if collect is not None: collect[f"{p}res"]=prev[:,:,::2,::2].mean(dim=(0,2,3))
        if mu is not None and f"{p}res" in mu: rb=rb-(mu[f"{p}res"][:,None]*(rq-rk)[0,0]).sum(0)
        res=F.conv2d(prev,rq.permute(3,2,0,1).contiguous(),rb,stride=2)
        prev=h(z+res)
    feat=torch.relu(sep(prev,"final/sep","final/bn","final"))
    B,C,H,Wd=feat.shape
    xx=feat.permute(0,2,3,1).reshape(B,H,Wd*C)
    xx=ps.bilstm_16(xx,W,"lstm1",f16,1,1); xx=ps.bilstm_16(xx,W,"lstm2",f16,1,1)
    xx=torch.relu(xx@T(W["dense1/kernel"])+T(W["dense1/bias"])); xx=no._bn(xx,W,"bn_dense",f32)
    return torch.sigmoid(xx@T(W["dense2/kernel"])+T(W["dense2/bias"])).numpy()
ref=no.forward(x,W)
def rep(name,o): e=np.abs(o-ref); print(f"{name:46s} max {e.max():.2e} mean {e.mean():.2e}",flush=True)
rep("fp16 (current)",fwd(x))
c_same={}; fwd(x,collect=c_same); rep("bias-corrected, mu from the same input",fwd(x,mu=c_same))
c_rand={}; fwd(xr,collect=c_rand); rep("bias-corrected, mu from uniform-random input",fwd(x,mu=c_rand))
c_oth={}; fwd(x2,collect=c_oth); rep("bias-corrected, mu from another recording",fwd(x,mu=c_oth))
