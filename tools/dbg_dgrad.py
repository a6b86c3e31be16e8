import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dnncancerannotator_b200 import native as N
N.lib()
n,h,w,ca,cb,co = 1,16,32,8,8,8
rng=np.random.default_rng(0)
bf=torch.bfloat16
cin=ca+cb
wt=(rng.normal(size=(3,3,cin,co))/np.sqrt(9*cin)).astype(np.float32)
dz=torch.from_numpy(rng.normal(size=(n,h,w,co)).astype(np.float32)).to(bf)
mask=torch.from_numpy(rng.normal(size=(n,h,w,ca)).astype(np.float32)).to(bf)
wd=torch.from_numpy(wt).cuda(); dzd=dz.cuda(); md=mask.cuda()
dx=torch.full((n,h,w,ca),3.0,dtype=bf,device='cuda'); dx2=torch.full((n,h,w,cb),3.0,dtype=bf,device='cuda')
dzv,mv,dxv,dx2v=(N.tensor_view(t) for t in (dzd,md,dx,dx2))
for use_mask in (True, False):
    N.call('dnnca_conv2d_dgrad', None, C.byref(dzv), N.ptr(wd), C.byref(dxv), C.byref(dx2v), 3, C.byref(mv) if use_mask else None, N.ACT_RELU if use_mask else N.ACT_NONE, 0.0, None, 0)
    torch.cuda.synchronize()
    xt=torch.zeros(n,cin,h,w,dtype=torch.float64,requires_grad=True)
    wtt=torch.from_numpy(wt).double().permute(3,2,0,1)
    y=torch.nn.functional.conv2d(xt,wtt,None,padding=1)
    y.backward(dz.double().permute(0,3,1,2))
    rdx=xt.grad.permute(0,2,3,1).numpy()
    ra=rdx[...,:ca]*((mask.float().numpy()>0) if use_mask else 1.0); rb=rdx[...,ca:]
    ea=np.abs(dx.float().cpu().numpy()-ra); eb=np.abs(dx2.float().cpu().numpy()-rb)
    print('mask',use_mask,'err a',ea.max(),'err b',eb.max())
    bad=np.argwhere(ea>0.05)
    print('bad a count',len(bad), bad[:10].tolist())
    bad=np.argwhere(eb>0.05)
    print('bad b count',len(bad), bad[:10].tolist())
