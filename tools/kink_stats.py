"""CPU experiment (precision modes): how many images a LeakyReLU-kink fallback of the fp16 path would have to flag.
Emulates the 16-bit path (fp16 operands at every stage, wide accumulation) on the canonical network and counts, per threshold eps,
the images with a hidden pre-activation |z| < eps.  Result (64 images): z1 std 3.55, fp16 error rms 1.5e-3 / max 5.9e-3;
eps = 6e-3 (4 sigma) flags 52 % of the images, eps = 1e-2 67 % -- a kink fallback costs as much as running everything at fp32 grade."""
import sys, time, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cnn as ocnn
import torch.nn.functional as F
cfg = ocnn.NetConfig.torch_flavour((256,256,1),2,[(32,3),(64,3)],[256,128],0.01)
p = ocnn.init_params(cfg, seed=7)
N=64
x = ocnn.synth_images(N,(256,256,1),seed=20251018)
t=time.time()
cache = ocnn.forward(cfg,p,x)
print("oracle fwd s", time.time()-t)
z1=cache.z[0].numpy(); z2=cache.z[1].numpy(); lg=cache.logits.numpy()
print("z1 std",z1.std(),"z2 std",z2.std(),"logit std",lg.std(), "margin std",(lg[:,0]-lg[:,1]).std())
# emulate fp16 path: fp16 operands, fp32 accumulate
def h(t): return t.half().float()
xt=torch.from_numpy(x).permute(0,3,1,2).float()
w0=torch.tensor(p.conv_w[0]).permute(0,3,1,2).float(); b0=torch.tensor(p.conv_b[0]).float()
w1=torch.tensor(p.conv_w[1]).permute(0,3,1,2).float(); b1=torch.tensor(p.conv_b[1]).float()
a0=F.max_pool2d(F.leaky_relu(F.conv2d(h(xt),h(w0),b0,padding=1),0.01),2)
a0=h(a0)
a1=F.leaky_relu(F.conv2d(a0,h(w1),b1,padding=1),0.01)
p1=h(F.max_pool2d(a1,2))
W1=torch.tensor(p.dense_w[0]).float()
flat=p1.reshape(N,-1)  # chw
z1h=(flat.double()@h(W1).double().T).numpy()+p.dense_b[0]
e1=z1h-z1
print("fp16 z1 err rms",e1.std(),"max",np.abs(e1).max())
h1=np.where(z1h>0,z1h,0.01*z1h)
z2h=h1@p.dense_w[1].T+p.dense_b[1]
e2=z2h-z2
print("fp16 z2 err rms",e2.std(),"max",np.abs(e2).max())
h2=np.where(z2h>0,z2h,0.01*z2h)
lgh=h2@p.dense_w[2].T+p.dense_b[2]
print("logit err max",np.abs(lgh-lg).max())
flips=((z1h>0)!=(z1>0)).any(1)|((z2h>0)!=(z2>0)).any(1)
print("actual flip images",flips.sum(),"of",N)
for eps in [1e-3,2e-3,4e-3,6e-3,1e-2]:
    fl=((np.abs(z1h)<eps).any(1)|(np.abs(z2h)<eps).any(1))
    print("eps",eps,"flagged frac",fl.mean(), "missed flips", (flips&~fl).sum())
