"""CPU experiment (precision modes): which fp16 rounding feeds the first dense pre-activation error.
Six sources (input, conv0 weights, pooled conv0 store, conv1 weights, pooled conv1 store, fc1 weights) each contribute
5-7e-4 rms of the 1.5e-3 total: no single stage can be made fp32-grade cheaply."""
import sys, time, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cnn as ocnn
import torch.nn.functional as F
torch.set_num_threads(8)
cfg = ocnn.NetConfig.torch_flavour((256,256,1),2,[(32,3),(64,3)],[256,128],0.01)
p = ocnn.init_params(cfg, seed=7)
N=16
x = ocnn.synth_images(N,(256,256,1),seed=20251018)
cache = ocnn.forward(cfg,p,x)
z1=cache.z[0].numpy()
def h(t): return t.half().double()
xt=torch.from_numpy(x).permute(0,3,1,2).double()
w0=torch.tensor(p.conv_w[0]).permute(0,3,1,2).double(); b0=torch.tensor(p.conv_b[0]).double()
w1=torch.tensor(p.conv_w[1]).permute(0,3,1,2).double(); b1=torch.tensor(p.conv_b[1]).double()
W1=torch.tensor(p.dense_w[0]).double()
def run(rx,rw0,rp1,rw1,rA,rW1):
    f=lambda t,r: h(t) if r else t
    a0=F.max_pool2d(F.leaky_relu(F.conv2d(f(xt,rx),f(w0,rw0),b0,padding=1),0.01),2)
    a0=f(a0,rp1)
    a1=F.leaky_relu(F.conv2d(a0,f(w1,rw1),b1,padding=1),0.01)
    pp=f(F.max_pool2d(a1,2),rA)
    z=(pp.reshape(N,-1)@f(W1,rW1).T).numpy()+p.dense_b[0]
    return z
names=["x","w0","P1","w1","A","W1"]
for i in range(6):
    flags=[False]*6; flags[i]=True
    e=run(*flags)-z1
    print(names[i],"rms",e.std(),"max",np.abs(e).max())
e=run(*[True]*6)-z1
print("all rms",e.std())
e=run(False,False,False,False,True,True)-z1
print("fc1 only (A,W1) rms",e.std())
e=run(True,True,True,True,False,False)-z1
print("conv only rms",e.std())
