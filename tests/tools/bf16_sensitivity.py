"""Which bf16 roundings move the gradient of configs/unet_big.yaml at random init? (CPU, oracle only)
   python tests/tools/bf16_sensitivity.py"""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, yaml
from oracle import ref_models as rm, ref_ops as ops
from dnncancerannotator_b200.synthetic import make_slices
def mk(round_fwd, round_bwd):
    class R(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x): return x.bfloat16().float() if round_fwd else x
        @staticmethod
        def backward(ctx, g): return g.bfloat16().float() if round_bwd else g
    return R.apply
def run(cfgname, size, B, mode):
    c = yaml.safe_load(open(f'/root/repo/configs/{cfgname}.yaml'))
    m = rm.build_model(c['model'], c['model_options'], (None,size,size,3), seed=3)
    x,y = make_slices(B,size,size,3,seed=77)
    oc, ot, ob, oa = ops.conv2d, ops.conv2d_transpose, ops.batchnorm, ops.activation
    rw = mode.get('w', False); ra = mk(mode.get('a', False), mode.get('g', False)); rbn = mk(mode.get('b', False), mode.get('g', False))
    def W(k): return (k + (k.detach().bfloat16().float() - k.detach())) if rw else k
    ops.conv2d = lambda x,k,b=None,padding='same',stride=1: oc(x, W(k) if k.shape[2]>=16 else k, b, padding, stride)
    ops.activation = lambda x,a: ra(oa(x,a)) if a is not None else x
    ops.conv2d_transpose = lambda x,k,b=None,stride=2: ra(ot(x, W(k), b, stride))
    def bn(*a, **kw):
        y, mm, mv = ob(*a, **kw); return rbn(y), mm, mv
    ops.batchnorm = bn
    try: r = m.train_step_grads(x, y, dict(weight_mul=3.0))
    finally: ops.conv2d, ops.conv2d_transpose, ops.batchnorm, ops.activation = oc, ot, ob, oa
    return m, r
def rel(a,b): return float(np.linalg.norm(a-b)/np.linalg.norm(b))
m, r0 = run('unet_big',64,2,{})
g0 = np.concatenate([r0['grads'][k].numpy().ravel() for k in m.trainable if not k.endswith('tconv/bias')])
for name, mode in [('weights only',dict(w=1)),('pre-BN acts (a) fwd only',dict(a=1)),('BN outputs (b) fwd only',dict(b=1)),('grads only',dict(g=1)),('all',dict(w=1,a=1,b=1,g=1))]:
    _, r1 = run('unet_big',64,2,mode)
    g1 = np.concatenate([r1['grads'][k].numpy().ravel() for k in m.trainable if not k.endswith('tconv/bias')])
    print(f'{name:28s} logits {rel(r1["logits"].numpy(), r0["logits"].numpy()):.4f}  grads {rel(g1,g0):.4f}')
