"""Rounding budget of a 16-bit storage pipeline (CPU, fp64 emulation; no GPU needed): the network's forward is evaluated in
fp64 twice -- exactly, and with a rounding to fp16 inserted at chosen SITES (the places where the dvae_b200 engine stores a
16-bit tensor or reads a 16-bit weight) -- and the relative L2 error of the outputs is printed per site, per group, and for
the combinations that were candidates for removal.  This is what decided the fp16 mode's design (DESIGN.md "Numerics"):
  x   input mel            wE/wD/wP  conv + LSTM weights (encoder / decoder / postnet)     w2  the small linears' weights
  yE/yD/yP  pre-BatchNorm convolution outputs      aE/aD/aP  BatchNorm+activation outputs   xgE/xgD  LSTM x-projections
  hE/hD  LSTM h per step    e  enc_linear output    z, d  decoder entry    rec  decoder output fed to the postnet
usage: python scripts/rounding_budget.py [rows_per_call=8]      (output of the run that drove the design:
profiles/r02_rounding_budget.txt)"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dvae_oracle as O
torch.set_num_threads(8)
R=int(sys.argv[1]) if len(sys.argv)>1 else 8
sd={k:(v.double() if v.is_floating_point() else v) for k,v in O.synth_state_dict(0).items()}
x1,x2,eps=O.synth_inputs(R)
eps=[e.double() for e in eps]
x=torch.cat([x1,x2]).double()
def rnd(t,on): return t.half().double() if on else t
def conv_bn_act(h,cp,bp,act,r,tag):
    w=r(sd[cp+".weight"],'w'+tag); b=sd[cp+".bias"]
    outs=[]
    for half in (h[:R],h[R:]):
        y=F.conv1d(half,w,b,padding=2)
        mean=y.mean((0,2),keepdim=True); var=y.var((0,2),unbiased=False,keepdim=True)
        y=r(y,'y'+tag)
        z=(y-mean)/torch.sqrt(var+1e-5)*sd[bp+".weight"].view(1,-1,1)+sd[bp+".bias"].view(1,-1,1)
        z=F.relu(z) if act=='relu' else (torch.tanh(z) if act=='tanh' else z)
        outs.append(r(z,'a'+tag))
    return torch.cat(outs)
def lstm(inp,prefix,layers,bi,r,tag):
    for layer in range(layers):
        outs_d=[]
        for suf in (("","_reverse") if bi else ("",)):
            w_ih=r(sd[f"{prefix}.weight_ih_l{layer}{suf}"],'w'+tag); w_hh=r(sd[f"{prefix}.weight_hh_l{layer}{suf}"],'w'+tag)
            b=sd[f"{prefix}.bias_ih_l{layer}{suf}"]+sd[f"{prefix}.bias_hh_l{layer}{suf}"]
            H=w_hh.shape[1]; T=inp.shape[1]
            xg=r(inp@w_ih.t()+b,'xg'+tag)
            hh=inp.new_zeros(inp.shape[0],H); c=inp.new_zeros(inp.shape[0],H); ys=[None]*T
            for t in (range(T-1,-1,-1) if suf else range(T)):
                gts=xg[:,t]+hh@w_hh.t()
                i_,f_,g_,o_=gts.split(H,1)
                c=torch.sigmoid(f_)*c+torch.sigmoid(i_)*torch.tanh(g_)
                hh=r(torch.sigmoid(o_)*torch.tanh(c),'h'+tag)
                ys[t]=hh
            outs_d.append(torch.stack(ys,1))
        inp=torch.cat(outs_d,-1)
    return inp
def fwd(S):
    r=lambda t,k: rnd(t, k in S)
    h=r(x,'x')
    for i in range(3): h=conv_bn_act(h,f"enc_modules.{i}.0.conv",f"enc_modules.{i}.1",'relu',r,'E')
    h=lstm(h.transpose(1,2),"enc_lstm",2,True,r,'E')
    flat=h.reshape(h.shape[0],-1)
    e=r(F.relu(F.linear(flat,r(sd["enc_linear.linear_layer.weight"],'w2'),sd["enc_linear.linear_layer.bias"])),'e')
    st=F.linear(e,r(sd["style.linear_layer.weight"],'w2'),sd["style.linear_layer.bias"])
    ct=F.linear(e,r(sd["content.linear_layer.weight"],'w2'),sd["content.linear_layer.bias"])
    smu=(st[:R,:4]+st[R:,:4])/2; slv=(st[:R,4:]+st[R:,4:])/2
    zs=eps[2]*torch.exp(0.5*slv)+smu
    z1=eps[0]*torch.exp(0.5*ct[:R,28:])+ct[:R,:28]; z2=eps[1]*torch.exp(0.5*ct[R:,28:])+ct[R:,:28]
    z=r(torch.cat([torch.cat([zs,z1],1),torch.cat([zs,z2],1)]),'z')
    d=r(F.linear(z,r(sd["dec_pre_linear1.weight"],'w2'),sd["dec_pre_linear1.bias"]),'d')
    d=r(F.linear(d,r(sd["dec_pre_linear2.weight"],'w2'),sd["dec_pre_linear2.bias"]),'d')
    h=lstm(d.view(2*R,64,128),"dec_lstm1",1,False,r,'D')
    h=h.transpose(1,2)
    for i in range(3): h=conv_bn_act(h,f"dec_modules.{i}.0",f"dec_modules.{i}.1",'relu',r,'D')
    h=lstm(h.transpose(1,2),"dec_lstm2",2,False,r,'D')
    rec=F.linear(h,r(sd["dec_linear2.linear_layer.weight"],'wD'),sd["dec_linear2.linear_layer.bias"]).transpose(1,2)
    p=r(rec,'rec')
    for i in range(5): p=conv_bn_act(p,f"postnet.convolutions.{i}.0.conv",f"postnet.convolutions.{i}.1",'tanh' if i<4 else None,r,'P')
    return {"slv":slv,"cmu":ct[:R,:28],"clv":ct[:R,28:],"rec":rec,"hat":rec+p}
ref=fwd(set())
ENC={'x','wE','w2','yE','aE','xgE','hE','e'}
DEC={'z','d','wD','yD','aD','xgD','hD'}
POST={'rec','wP','yP','aP'}
ALL=ENC|DEC|POST
def rep(name,S):
    o=fwd(S)
    print(f"{name:36s}"+" ".join(f"{k}={((o[k]-ref[k]).norm()/ref[k].norm()).item():.2e}" for k in ref),flush=True)
rep("all roundings",ALL)
rep("encoder only",ENC)
rep("decoder only",DEC)
rep("postnet only",POST)
for k in sorted(DEC|POST): rep(f"only {k}",{k})
rep("all but x,w2,e",ALL-{'x','w2','e'})
rep("all but x,w2,e,yE,yD,yP,xgE,xgD",ALL-{'x','w2','e','yE','yD','yP','xgE','xgD'})
rep("all but x,w2,e,y*,xg*,w*",ALL-{'x','w2','e','yE','yD','yP','xgE','xgD','wE','wD','wP'})
rep("all but x,w2,e,wE,wD",ALL-{'x','w2','e','wE','wD'})
