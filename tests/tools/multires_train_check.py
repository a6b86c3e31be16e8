"""MultiResUnet training parity on the GPU box: one training step (forward with batch statistics, weighted BCE,
full backward) of the CUDA path against the torch-CPU oracle (oracle/ref_models.py, autograd), per variable.

    python tests/tools/multires_train_check.py [--size 32] [--batch 2] [--modes fp32,bf16] [--steps 0]

Writes gpurun_out/multires_train_check.json.  Test infrastructure (imports the oracle)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref_models as rm          # noqa: E402


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=32)
    ap.add_argument('--batch', type=int, default=2)
    ap.add_argument('--modes', default='fp32,bf16')
    ap.add_argument('--graph', type=int, default=1)
    ap.add_argument('--seed', type=int, default=11)
    ap.add_argument('--out', default='gpurun_out/multires_train_check.json')
    args = ap.parse_args()
    from dnncancerannotator_b200.models import tf_models
    from dnncancerannotator_b200.synthetic import make_slices
    B, S = args.batch, args.size
    x, y = make_slices(B, S, S, 5, seed=args.seed)
    ref = rm.build_model('MultiResUnet', dict(height=None, width=None, n_channels=5), None, seed=0)
    ref.randomize_bn(seed=1)
    from oracle import ref_ops
    rec, orig = [], ref_ops.maxpool

    def spy(t, rate=2, return_indices=False):
        out, idx = orig(t, rate, return_indices=True)
        rec.append((idx.numpy(), t.detach().numpy()))
        return (out, idx) if return_indices else out
    ref_ops.maxpool = spy
    want = ref.train_step_grads(x, y, dict(weight_mul=3.0))
    ref_ops.maxpool = orig
    report = {}
    for mode in args.modes.split(','):
        m = tf_models.MultiResUnet(None, None, 5, dtype=mode)
        m.use_cuda_graph = bool(args.graph)
        m.build((None, S, S, 5))
        m.set_weights(ref.get_weights())
        m.compile(loss=dict(class_name='WeightedCrossentropy', config=dict(weight_mul=3.0)))
        for _ in range(4 if args.graph else 1):
            ps = m.forward_backward(x, y).cpu().numpy()
        logits = m.last_logits.cpu().numpy()
        g = m.get_grads()
        wl = want['logits'].numpy()
        rows = {k: rel_l2(g[k], want['grads'][k].numpy()) for k in g}
        allg = np.concatenate([g[k].ravel() for k in g])
        allw = np.concatenate([want['grads'][k].numpy().ravel() for k in g])
        st = m.get_weights()
        mov = {k: rel_l2(st[k], want['new_moving'][k].numpy()) for k in want['new_moving']}
        worst = sorted(rows.items(), key=lambda kv: -kv[1])[:12]
        rep = dict(logits_rel_l2=rel_l2(logits, wl), logits_rel_max=float(np.abs(logits - wl).max() / np.abs(wl).max()),
                   loss=float(ps.mean()), loss_ref=want['data_loss'], grads_rel_l2=rel_l2(allg, allw),
                   grads_nonfinite=int((~np.isfinite(allg)).sum()), worst=worst,
                   moving_worst=sorted(mov.items(), key=lambda kv: -kv[1])[:4],
                   first_layers={k: rows[k] for k in list(rows)[:10]}, last_layers={k: rows[k] for k in list(rows)[-8:]})
        # max-pool argmax of the training plan vs the oracle's (logical channels of the padded block outputs)
        from dnncancerannotator_b200 import runtime as R
        plan = m.training_plan(B, S, S)
        pools = [op for op in plan.ops if isinstance(op, R.PoolOp)]
        pm = []
        for lvl, op in enumerate(pools):
            W = 1.67 * 32 * 2 ** lvl
            f = [int(W * 0.167), int(W * 0.333), int(W * 0.5)]
            pp = [(c + 15) // 16 * 16 for c in f]
            pos = np.concatenate([np.arange(f[0]), pp[0] + np.arange(f[1]), pp[0] + pp[1] + np.arange(f[2])])
            ours = op.idx.cpu().numpy()[..., pos]
            oidx, oin = rec[lvl]
            diff = ours != oidx
            xin = op.x.torch_view().float().cpu().numpy()[..., pos]
            pm.append(dict(level=lvl, mismatches=int(diff.sum()), of=int(diff.size),
                           input_rel_l2=rel_l2(xin, oin), input_max_abs=float(np.abs(xin - oin).max())))
        rep['pool_argmax'] = pm
        rep['all'] = rows
        report[mode] = rep
        print(mode, json.dumps({k: v for k, v in rep.items() if k != 'all'}, indent=1))
    os.makedirs('gpurun_out', exist_ok=True)
    with open(args.out, 'w') as f:
        json.dump(report, f, indent=1)


if __name__ == '__main__':
    main()
