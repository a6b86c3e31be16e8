"""Debug: run-to-run determinism of the BN-folded plan (which buffer diverges first, eager vs graph)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_models as rm
from dnncancerannotator_b200.models import tf_models
from dnncancerannotator_b200.synthetic import make_slices
from dnncancerannotator_b200 import runtime as R

OPTS = dict(n_filters_first=32, n_downsample=2, rate=2, kernel_size=3, conv_stride=1, bn=True, padding='same')
ref = rm.build_model('UNetAnnotator', OPTS, (None, 64, 64, 3), seed=4)
ref.randomize_bn(seed=6)
x, y = make_slices(4, 64, 64, 3, seed=9)
for fold in (1, 0):
    os.environ['DNNCA_BN_FOLD'] = str(fold)
    m = tf_models.UNetAnnotator(**OPTS, dtype='bf16')
    m.build((None, 64, 64, 3))
    m.compile(loss=dict(class_name='WeightedCrossentropy', config=dict(weight_mul=3.0)))
    m.set_weights(ref.get_weights())
    snaps = []
    for rep in range(8):
        m.forward_backward(x, y)
        torch.cuda.synchronize()
        plan = m._plan(4, 64, 64)
        snap = {b.name + f'#{i}': b.data.float().clone() for i, b in enumerate(plan.bufs) if b.data is not None}
        snap['logits'] = plan.logits.clone()
        for i, op in enumerate(plan.ops):
            if isinstance(op, R.BNOp):
                snap[f'bn{i}.ss'] = op.ss.clone()
            if getattr(op, 'scratch', None) is not None:
                snap[f'op{i}.scratch'] = op.scratch.clone()
        snap['stats'] = plan.stats.clone().float()
        g = m.get_grads()
        snap['grads'] = torch.tensor(np.concatenate([v.ravel() for v in g.values()]))
        snaps.append(snap)
    print('fold', fold, 'bufs', len(snaps[0]))
    for rep in range(1, 8):
        bad = []
        for k in snaps[0]:
            a, b = snaps[rep][k].double().cpu(), snaps[0][k].double().cpu()
            d = float((a - b).norm() / max(float(b.norm()), 1e-30))
            if d > 0:
                bad.append((k, d))
        print(' rep', rep, 'first diffs:', [(k, f'{d:.2e}') for k, d in bad[:6]], 'n_diff', len(bad))
