"""Parity of the bf16 CUDA path against the fp32 oracle at the BASELINE configs' REAL shapes and bench conditions
(VERDICT r1 item 1a): 256x256 slices, per-GPU batch >= 8, glorot init seed 0 with gamma = 1 / beta = 0, and again after
N fp32-oracle Adam steps ("conditioned weights").  For every point it also runs the oracle's own bf16-storage emulation
(oracle/ref_bf16.py) so that the table separates what the implementation adds from what bf16 storage costs by itself.

The fp32 oracle is evaluated with torch ON THE GPU BOX'S DEVICE with TF32 disabled (same restatement, same code; the
CPU evaluation of it is compared at a small size first and reported as `oracle_device_vs_cpu`), because 20 Adam steps of
unet_big at 256^2 x 8 cost minutes on the host cores.  Test infrastructure only: nothing here is on the product path.

  python tests/tools/parity_real_shapes.py --out gpurun_out/parity_real_shapes.json [--size 256 --batch 8 --steps 20]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_models as rm, ref_ops as ops                  # noqa: E402
from oracle.ref_bf16 import emulate_bf16                             # noqa: E402
from dnncancerannotator_b200.models import tf_models                 # noqa: E402
from dnncancerannotator_b200.synthetic import make_slices           # noqa: E402
from dnncancerannotator_b200.utils.load import load_config          # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def relmax(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30))


def cat(d, names):
    return np.concatenate([np.asarray(d[k].detach().cpu().numpy() if torch.is_tensor(d[k]) else d[k]).ravel() for k in names])


def point(ref, m, x, y, loss_cfg, tag):
    names = [k for k in ref.trainable if not k.endswith('/tconv/bias')]
    r = ref.train_step_grads(x, y, loss_cfg)
    with emulate_bf16():
        e = ref.train_step_grads(torch.tensor(x).bfloat16().float(), y, loss_cfg)
    m.set_weights(ref.get_weights())
    m.use_cuda_graph = False
    per = m.forward_backward(x, y).cpu().numpy()
    logits = m.last_logits.cpu().numpy()
    g = m.get_grads()
    rl = r['logits'].cpu().numpy()
    allr, alle, allg = cat(r['grads'], names), cat(e['grads'], names), cat(g, names)
    out = dict(tag=tag,
               ours=dict(logits_rel_l2=rel(logits, rl), logits_rel_max=relmax(logits, rl),
                         loss_rel=abs(float(per.mean()) - r['data_loss']) / abs(r['data_loss']),
                         grad_rel_l2=rel(allg, allr), grad_cosine=float(allg @ allr / (np.linalg.norm(allg) * np.linalg.norm(allr)))),
               emulated_bf16_storage=dict(logits_rel_l2=rel(e['logits'].cpu().numpy(), rl),
                                          logits_rel_max=relmax(e['logits'].cpu().numpy(), rl),
                                          loss_rel=abs(e['data_loss'] - r['data_loss']) / abs(r['data_loss']),
                                          grad_rel_l2=rel(alle, allr)),
               ours_vs_emulation=dict(logits_rel_l2=rel(logits, e['logits'].cpu().numpy()), grad_rel_l2=rel(allg, alle)),
               masks_disagree_p05=float(((logits > 0) != (rl > 0)).mean()),
               masks_disagree_p08=float(((logits >= np.log(4.0)) != (rl >= np.log(4.0))).mean()),
               data_loss=r['data_loss'])
    # per-layer-group gradient error (which part of the net carries it)
    groups = {}
    for k in names:
        gname = k.split('/')[0] + '/' + k.split('/')[1] if k != 'head/kernel' and k != 'head/bias' else 'head'
        groups.setdefault(gname, []).append(k)
    out['grad_rel_l2_by_block'] = {gn: rel(cat(g, ks), cat(r['grads'], ks)) for gn, ks in groups.items()}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=256)
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--configs', default='unet_big,mulmo_unet,unet')
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'parity_real_shapes.json'))
    a = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device('cuda', 0)
    res = dict(size=a.size, batch=a.batch, adam_steps=a.steps, tolerances=dict(logits=1e-2, loss=1e-3, grad=2e-2), configs={})
    for name in a.configs.split(','):
        cfg = load_config([os.path.join(ROOT, 'configs', name + '.yaml'),
                           os.path.join(ROOT, 'configs', 'additionals', 'deploy_options.yaml')])
        loss_cfg = cfg['deploy_options']['loss']['config']
        t0 = time.time()
        # the oracle on the device == the oracle on the CPU (small size, same weights)
        small = rm.build_model(cfg['model'], cfg['model_options'], (None, 64, 64, 3), seed=0)
        xs, ys = make_slices(2, 64, 64, 3, seed=5)
        rc = small.train_step_grads(xs, ys, loss_cfg)
        small.to(dev)
        rd = small.train_step_grads(xs, ys, loss_cfg)
        nm = [k for k in small.trainable if not k.endswith('/tconv/bias')]
        dev_vs_cpu = dict(logits_rel_l2=rel(rd['logits'].cpu().numpy(), rc['logits'].numpy()),
                          grad_rel_l2=rel(cat(rd['grads'], nm), cat(rc['grads'], nm)))
        ref = rm.build_model(cfg['model'], cfg['model_options'], (None, a.size, a.size, 3), seed=0).to(dev)
        m = getattr(tf_models, cfg['model'])(**cfg['model_options'], dtype='bf16')
        m.build((None, a.size, a.size, 3))
        m.compile(loss=cfg['deploy_options']['loss'])
        x, y = make_slices(a.batch, a.size, a.size, 3, seed=1234)
        pts = [point(ref, m, x, y, loss_cfg, 'glorot init, gamma=1 beta=0 (bench conditions)')]
        mom = {k: (torch.zeros_like(ref.weights[k]), torch.zeros_like(ref.weights[k])) for k in ref.trainable}
        losses = []
        for t in range(a.steps):
            r = ref.train_step_grads(x, y, loss_cfg)
            losses.append(r['loss'])
            for k in ref.trainable:
                ref.weights[k], m_, v_ = ops.adam_step(ref.weights[k], r['grads'][k], mom[k][0], mom[k][1], t + 1)
                mom[k] = (m_, v_)
            for k, v in r['new_moving'].items():
                ref.weights[k] = v
        pts.append(point(ref, m, x, y, loss_cfg, f'after {a.steps} fp32-oracle Adam steps'))
        res['configs'][name] = dict(points=pts, oracle_device_vs_cpu=dev_vs_cpu, oracle_losses=[losses[0], losses[-1]],
                                    seconds=round(time.time() - t0, 1))
        print(name, json.dumps(res['configs'][name])[:600], flush=True)
        del m, ref
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    with open(a.out, 'w') as f:
        json.dump(res, f, indent=1)
    print('wrote', a.out)


if __name__ == '__main__':
    main()
