"""Generates the committed golden vectors under tests/golden/ from the oracle.

The reference cannot run here (TensorFlow is absent) and ships no
vectors of its own, so these fixtures pin the *oracle restatement* (seeded
inputs, weights, logits, loss, all parameter gradients, pool indices).  They let
the GPU tests compare against committed numbers and make any later change to the
oracle visible.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_models as rm                      # noqa: E402
from dnncancerannotator_b200.synthetic import make_slices  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (model, model_options, B, H, C, loss_config, randomize_bn)
    'unet_tiny': ('UNetAnnotator', dict(n_filters_first=3, n_downsample=2, rate=2, kernel_size=3, conv_stride=1,
                                        bn=False, padding='same'), 2, 32, 3, dict(weight_mul=3.0), False),
    'unet_bn_tiny': ('UNetAnnotator', dict(n_filters_first=4, n_downsample=2, rate=2, kernel_size=3, conv_stride=1,
                                           bn=True, padding='same'), 3, 32, 3, dict(weight_mul=3.0), True),
    'mulmo_tiny': ('MulmoUNetAnnotator', dict(n_filters_first=4, n_downsample=2, rate=2, kernel_size=3,
                                              conv_stride=1, bn=True, padding='same'), 2, 32, 3,
                   dict(weight_mul=3.0), True),
    'unet_leaky_l2_tiny': ('UNetAnnotator', dict(n_filters_first=3, n_downsample=2, rate=2, kernel_size=3,
                                                 conv_stride=1, bn=False, padding='same',
                                                 activation=dict(class_name='LeakyReLU', config=dict(alpha=0.3)),
                                                 kernel_regularizer=dict(class_name='L2', config=dict(l2=0.01))),
                           2, 32, 5, dict(weight_mul=3.0), False),
}


def multires_train_summary(r, trainable, grads=None):
    """The fixture's view of a MultiResUnet training step: ``r`` = oracle ``train_step_grads`` result, or (for the CUDA
    path) a dict with logits / loss / per_sample / new_moving and ``grads`` = name -> array."""
    g = grads if grads is not None else {k: v.numpy() for k, v in r['grads'].items()}
    out = dict(loss=np.float32(r['loss']), logits=np.asarray(r['logits'], np.float32), per_sample=np.asarray(r['per_sample'], np.float32))
    for k in trainable:
        if k.endswith('/kernel'):
            out['n:' + k] = np.float32(np.linalg.norm(g[k].astype(np.float64)))
        else:
            out['g:' + k] = g[k]
    out['sample'] = np.concatenate([g[k].ravel() for k in trainable])[::997].copy()
    for k, v in r['new_moving'].items():
        out['m:' + k] = np.asarray(v, np.float32)
    return out


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    for name, (model, opts, B, H, C, loss_cfg, rbn) in CASES.items():
        m = rm.build_model(model, opts, (None, H, H, C), seed=0)
        if rbn:
            m.randomize_bn(seed=1)
        else:
            rng = np.random.default_rng(1)
            for k in m.weights:
                if k.endswith('/bias'):
                    m.weights[k] = torch.tensor(rng.normal(0, 0.05, m.weights[k].shape), dtype=torch.float32)
        x, y = make_slices(B, H, H, C, seed=1234)
        r = m.train_step_grads(x, y, loss_cfg)
        ev = m.forward(x, training=False)
        out = dict(x=x, y=y, loss=np.float32(r['loss']), data_loss=np.float32(r['data_loss']),
                   per_sample=r['per_sample'].numpy(), logits=r['logits'].numpy(),
                   eval_logits=ev['logits'].detach().numpy())
        for k, v in m.get_weights().items():
            out['w:' + k] = v
        for k, v in r['grads'].items():
            out['g:' + k] = v.numpy()
        for k, v in r['new_moving'].items():
            out['m:' + k] = v.numpy()
        for i, idx in enumerate(r['pool_idx']):
            out[f'pool_idx:{i}'] = idx.numpy()
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
        print(name, 'loss', r['loss'], 'params', m.n_params())
    # MultiResUnet forward (inference mode, BN moving stats): weights are re-derived from seed 0
    m = rm.build_model('MultiResUnet', dict(height=None, width=None, n_channels=5), None, seed=0)
    m.randomize_bn(seed=1)
    x, _ = make_slices(1, 32, 32, 5, seed=1234)
    ev = m.forward(x, training=False)
    np.savez_compressed(os.path.join(HERE, 'multires_fwd_tiny.npz'), x=x, eval_logits=ev['logits'].detach().numpy())
    print('multires logits', float(ev['logits'].abs().mean()))
    # MultiResUnet TRAINING step (batch statistics, weighted BCE, all gradients).  7.2 M gradient values do not belong in a
    # fixture: kept are the logits, the loss, every BatchNorm gradient and moving statistic in full, the L2 norm of every
    # kernel gradient and every 997th element of the concatenated gradient.  Batch seed 1235: the oracle agrees with its
    # own fp64 evaluation to 1.4e-5 there (no relu / max-pool near-ties).
    x, y = make_slices(2, 32, 32, 5, seed=1235)
    r = m.train_step_grads(x, y, dict(weight_mul=3.0))
    out = multires_train_summary(r, m.trainable)
    out.update(x=x, y=y)
    np.savez_compressed(os.path.join(HERE, 'multires_train_tiny.npz'), **out)
    print('multires train loss', r['loss'])


if __name__ == '__main__':
    main()
