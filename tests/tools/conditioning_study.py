"""How well-conditioned are the BatchNorm configs (unet_big.yaml, mulmo_unet.yaml) as functions of their stored
activations and weights?  Oracle only (torch-CPU), no GPU, no product code.

For each config the fp32 oracle step (forward + weighted BCE + backward) is compared with the SAME oracle whose stored
tensors (activations after every conv / BN, their gradients, and the weights of the tensor-core layers) are rounded to
m mantissa bits, m in {7 (bfloat16), 10 (TF32 operands / fp16), 13, 16}.  If the deviation scales like 2^-m with one
constant, the constant is the network's own amplification of storage rounding -- a property of the function at these
weights, independent of who implements it -- and the table says how many mantissa bits ANY implementation needs to meet
the north-star tolerances (logits 1e-2, gradients 2e-2).  m = 10 is what the reference's own default GPU path (TensorFlow
>= 2.4 on Ampere: TF32 convolutions) computes with.

Also after 20 fp32 Adam steps ("conditioned weights").

  python tests/tools/conditioning_study.py [--size 128 --batch 8 --steps 20] > profiles/r02_conditioning.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_models as rm, ref_ops as ops                  # noqa: E402
from oracle.ref_bf16 import emulate_bf16, round_mantissa             # noqa: E402
from dnncancerannotator_b200.synthetic import make_slices           # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def flat_grads(m, r):
    names = [k for k in m.trainable if not k.endswith('tconv/bias')]     # exactly zero (ConvT -> BN), rounding noise only
    return np.concatenate([r['grads'][k].numpy().ravel() for k in names])


def compare(m, x, y, loss_cfg, bits_list):
    r0 = m.train_step_grads(x, y, loss_cfg)
    g0 = flat_grads(m, r0)
    rows = {}
    for bits in bits_list:
        with emulate_bf16(mantissa_bits=bits):
            r1 = m.train_step_grads(round_mantissa(torch.tensor(x), bits), y, loss_cfg)
        g1 = flat_grads(m, r1)
        rows[str(bits)] = dict(logits_rel_l2=rel(r1['logits'].numpy(), r0['logits'].numpy()), grad_rel_l2=rel(g1, g0),
                               grad_cosine=float(g1 @ g0 / (np.linalg.norm(g1) * np.linalg.norm(g0))),
                               loss_rel=abs(r1['data_loss'] - r0['data_loss']) / abs(r0['data_loss']),
                               amplification_logits=rel(r1['logits'].numpy(), r0['logits'].numpy()) / 2.0 ** -(bits + 1),
                               amplification_grads=rel(g1, g0) / 2.0 ** -(bits + 1))
    return rows, r0


def adam_steps(m, x, y, loss_cfg, n):
    mom = {k: (torch.zeros_like(m.weights[k]), torch.zeros_like(m.weights[k])) for k in m.trainable}
    losses = []
    for t in range(n):
        r = m.train_step_grads(x, y, loss_cfg)
        losses.append(r['loss'])
        for k in m.trainable:
            m.weights[k], a, b = ops.adam_step(m.weights[k], r['grads'][k], mom[k][0], mom[k][1], t + 1)
            mom[k] = (a, b)
        for k, v in r['new_moving'].items():
            m.weights[k] = v
    return losses


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=128)
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--configs', default='unet_big,mulmo_unet,unet')
    a = ap.parse_args()
    bits_list = [7, 10, 13, 16]
    loss_cfg = dict(weight_mul=3.0)
    out = dict(note=__doc__.split('\n\n')[1].replace('\n', ' '), size=a.size, batch=a.batch, bits=bits_list, configs={})
    for name in a.configs.split(','):
        c = yaml.safe_load(open(os.path.join(ROOT, 'configs', name + '.yaml')))
        m = rm.build_model(c['model'], c['model_options'], (None, a.size, a.size, 3), seed=3)
        x, y = make_slices(a.batch, a.size, a.size, 3, seed=77)
        t0 = time.time()
        init, _ = compare(m, x, y, loss_cfg, bits_list)
        losses = adam_steps(m, x, y, loss_cfg, a.steps)
        after, _ = compare(m, x, y, loss_cfg, bits_list)
        out['configs'][name] = dict(at_init=init, after_adam_steps=after, adam_steps=a.steps,
                                    loss_first=losses[0], loss_last=losses[-1], seconds=round(time.time() - t0, 1))
        print(name, json.dumps(out['configs'][name])[:400], file=sys.stderr, flush=True)
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
