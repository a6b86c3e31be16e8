#!/usr/bin/env python
"""One eager training step of a config between cudaProfilerStart/Stop (for `ncu --profile-from-start off`).

  python tools/ncu_step.py --config mulmo_unet --batch 8
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ['DNNCA_NO_GRAPH'] = '1'
os.environ.setdefault('DNNCA_WGRAD_STREAM', '0')       # serial launch order: one kernel at a time under the profiler
os.environ.setdefault('DNNCA_BRANCH_STREAMS', '0')
from dnncancerannotator_b200.models import tf_models                 # noqa: E402
from dnncancerannotator_b200.synthetic import make_slices           # noqa: E402
from dnncancerannotator_b200.utils.load import load_config          # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--config', default='mulmo_unet')
ap.add_argument('--batch', type=int, default=8)
ap.add_argument('--size', type=int, default=256)
a = ap.parse_args()
cfg = load_config([os.path.join(ROOT, 'configs', a.config + '.yaml'), os.path.join(ROOT, 'configs', 'additionals', 'deploy_options.yaml')])
m = getattr(tf_models, cfg['model'])(**cfg['model_options'], dtype='bf16')
m.build((None, a.size, a.size, 3))
m.compile(optimizer=cfg['deploy_options']['optimizer'], loss=cfg['deploy_options']['loss'])
x, y = make_slices(a.batch, a.size, a.size, 3, seed=1)
for _ in range(2):
    m.train_step(x, y)
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = m.train_step(x, y)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('loss', float(loss))
