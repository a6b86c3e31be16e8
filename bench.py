#!/usr/bin/env python
"""bench.py -- train slices/s of configs/unet.yaml on N B200s (BASELINE.json metric).

  python bench.py --gpus 1 --steps K --warmup W            # this build (CUDA path)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...                      # oracle port of the reference on host cores

A "step" = one optimizer step (zero grads, forward, fused head+loss, backward, gradient all-reduce
when N>1, fused Adam) of UNetAnnotator(configs/unet.yaml) on a per-GPU batch of synthetic
256x256x3 slices.  ``value`` is measured with the batch already resident in HBM (CUDA-graph replay,
CUDA events, max over ranks); ``e2e`` goes through the public ``Model.train_step`` with pinned HOST
buffers (H2D of the inputs and D2H of the loss inside the timed region).  One JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'train_slices_per_sec'
UNIT = 'slices/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='unet')
    ap.add_argument('--batch', type=int, default=256, help='per-GPU batch (weak scaling)')
    ap.add_argument('--size', type=int, default=256)
    ap.add_argument('--channels', type=int, default=3)
    ap.add_argument('--dtype', default='bf16')
    ap.add_argument('--cpu-batch', type=int, default=32,
                    help='batch of the CPU arm (measured on the B200 host, 16 threads: 181 slices/s at 4, 318 at 8, 405 at 16, 492 at 32)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-profile', action='store_true')
    ap.add_argument('--forward', action='store_true',
                    help='forward-only (inference) batch sweep: --config multiresunet --forward --batches 1,2,...')
    ap.add_argument('--batches', default='1,2,4,8,16,32,64,128,256')
    ap.add_argument('--no-secondary', dest='secondary', action='store_false',
                    help='skip the short unet_big / mulmo_unet runs appended to the default (unet.yaml) line')
    ap.add_argument('--no-f32-e2e', action='store_true', help='skip the float32-input variant of the e2e loop')
    ap.add_argument('--no-wc', dest='wc', action='store_false',
                    help='plain pinned host buffers instead of write-combined ones for the e2e loop')
    return ap.parse_args()


def load_cfg(name):
    from dnncancerannotator_b200.utils.load import load_config
    return load_config([os.path.join(ROOT, 'configs', name + '.yaml'),
                        os.path.join(ROOT, 'configs', 'additionals', 'deploy_options.yaml')])


# ---------------------------------------------------------------------------------------------
# CPU comparator: the oracle port of the reference (TensorFlow itself is not installable here)
# ---------------------------------------------------------------------------------------------
def cpu_reference_rate(cfg, args, steps, warmup, budget_s=25.0):
    """torch-CPU restatement of the reference train step (forward + weighted BCE + backward +
    keras Adam) on all host threads; a bounded sample of the same workload (batch ``cpu_batch``)."""
    import torch
    from oracle import ref_models as rm
    from oracle import ref_ops as ops
    from dnncancerannotator_b200.synthetic import make_slices
    # all the host threads this process may use: torchrun exports OMP_NUM_THREADS=1 to every rank, which would leave
    # the CPU arm single-threaded at N > 1 (the round-1 SCALE ratios); the arm runs on rank 0 alone, the other ranks idle
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    torch.set_num_threads(max(ncpu, 1))
    B = args.cpu_batch
    ref = rm.build_model(cfg['model'], cfg['model_options'], (None, args.size, args.size, args.channels), seed=0)
    x, y = make_slices(B, args.size, args.size, args.channels, seed=1234)
    loss_cfg = cfg['deploy_options']['loss']['config']
    mom = {k: (torch.zeros_like(ref.weights[k]), torch.zeros_like(ref.weights[k])) for k in ref.trainable}

    def step(t):
        r = ref.train_step_grads(x, y, loss_cfg)
        for k in ref.trainable:
            ref.weights[k], m_, v_ = ops.adam_step(ref.weights[k], r['grads'][k], mom[k][0], mom[k][1], t + 1)
            mom[k] = (m_, v_)
        for k, v in r['new_moving'].items():
            ref.weights[k] = v
    for t in range(warmup):
        step(t)
    times = []
    t_begin = time.perf_counter()
    for t in range(steps):
        t0 = time.perf_counter()
        step(warmup + t)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > budget_s:
            break
    ms = 1e3 * statistics.median(times)
    return dict(value=B / (ms / 1e3), ms_per_step=ms, steps=len(times), cores=torch.get_num_threads(),
                sample=f'{len(times)} steps of batch {B} ({args.size}x{args.size}x{args.channels} fp32), median')


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cfg = load_cfg(args.config)
    r = cpu_reference_rate(cfg, args, steps=args.steps, warmup=max(args.warmup, 1), budget_s=120.0)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': r['value'], 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': r['steps'], 'warmup': args.warmup, 'ms_per_step': r['ms_per_step'],
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'configs/{args.config}.yaml {cfg["model"]} training step, per-GPU batch {args.batch} of '
                               f'{args.size}x{args.size}x{args.channels} slices ({args.dtype} activations, fp32 accumulate/master weights)',
                   'global_batch': args.batch * args.gpus, 'parallelism': f'dp{args.gpus}',
                   'sample': f'each timed step = batch {args.cpu_batch} of the same slices on the host CPU (fp32)'},
        'cpu_baseline': {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': 'port',
                         'sample': r['sample'] + '; torch-CPU restatement of the reference (TensorFlow unavailable)'},
        'e2e': {'value': r['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
            except (ValueError, IndexError):
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        # under load = the upper half of the samples (idle samples before/after the region are excluded)
        sm.sort()
        load = sm[len(sm) // 2:] if sm else []
        return {'sm_mhz': statistics.median(load) if load else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


def profile_step(m, plan, cfg, world):
    """Per-kernel CUDA-event pass (eager, after the timed region): launches per step, breakdown, categories and the
    roofline of the dominant kernel -- for the wide (tensor-bound) configs the dominant CONV kernel."""
    import torch
    from dnncancerannotator_b200 import native as N
    lib = N.lib()
    m.use_cuda_graph = False
    lib.dnnca_debug_launch_count(1)
    m._train_on_static(plan)
    torch.cuda.synchronize()
    launches_per_step = int(lib.dnnca_debug_launch_count(1))
    # per-kernel times are taken with every kernel ALONE on the device: the step graph runs weight gradients and
    # MulmoUNet's encoder branches on side streams, where the events around a call would time it sharing the SMs
    saved_env = {k: os.environ.get(k) for k in ('DNNCA_WGRAD_STREAM', 'DNNCA_BRANCH_STREAMS')}
    os.environ['DNNCA_WGRAD_STREAM'] = '0'
    os.environ['DNNCA_BRANCH_STREAMS'] = '0'
    peaks = {}
    pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    # every iteration is queued behind a device-side delay long enough for the host to enqueue the whole step first: the
    # events then bracket kernel execution alone (a Python launch costs ~50 us -- about one of these kernels -- and with an
    # empty queue the gap between an event record and the launch behind it would be counted into the kernel)
    delay_cycles = int(min(max(launches_per_step * 60e-6, 3e-3), 60e-3) * 2.0e9)
    with N.Profiler() as prof:
        for _ in range(3):
            torch.cuda._sleep(delay_cycles)
            m._train_on_static(plan)
    agg = prof.summary()
    for k, v in saved_env.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    m.use_cuda_graph = True
    tot = sum(d['ms'] for d in agg.values())
    rows = sorted(agg.items(), key=lambda kv: -kv[1]['ms'])
    breakdown = [{'kernel': k, 'share': round(d['ms'] / tot, 4), 'ms': round(d['ms'] / 3, 4),
                  'gbs': round(d['bytes'] / d['ms'] / 1e6, 1) if d['ms'] else None,
                  'tflops': round(d['flops'] / d['ms'] / 1e9, 2) if d['ms'] else None} for k, d in rows[:40]]
    cats = {}
    for kname, v in agg.items():
        base = kname.split('[')[0]
        cat = ('conv_' + base.split('_')[-1] if base.startswith('conv2d') else
               'convT' if base.startswith('convtranspose') else
               'batchnorm' if base.startswith('bn_') or base == 'channel_stats' else
               'pool' if base.startswith('maxpool') else
               'head_loss' if base.startswith('head') or base.startswith('label') or base.startswith('loss') else base)
        c = cats.setdefault(cat, [0.0, 0.0, 0.0])
        c[0] += v['ms'] / 3
        c[1] += v['flops'] / 3
        c[2] += v['bytes'] / 3
    categories = {k2: {'ms': round(v[0], 3), 'share': round(v[0] * 3 / tot, 4),
                       'tflops': round(v[1] / v[0] / 1e9, 1) if v[1] else None,
                       'gbs': round(v[2] / v[0] / 1e6, 1)} for k2, v in sorted(cats.items(), key=lambda kv: -kv[1][0])}
    tens_sus = float(peaks.get('bf16_tflops_sustained', 1400.0))      # kernels timed inside a long step
    tens_burst = float(peaks.get('bf16_tflops', 1590.0))
    conv_ms = sum(v['ms'] for v in agg.values() if v['flops'])
    conv_fl = sum(v['flops'] for v in agg.values())
    src = 'measured (MEASURED_PEAKS.json)' if peaks else 'fallback (B200_PROFILING.md)'
    wide = cfg['model_options'].get('n_filters_first', 64) >= 16
    if wide:
        conv_rows = [(k, d) for k, d in rows if d['flops']]
        k, d = conv_rows[0] if conv_rows else rows[0]
    else:
        k, d = rows[0]
    if wide and d['flops']:
        ach = d['flops'] / d['ms'] / 1e9
        roofline = {'bound': 'tensor', 'kernel': k, 'achieved': round(ach, 1), 'peak': tens_sus, 'unit': 'TFLOP/s',
                    'frac': round(ach / tens_sus, 4), 'traffic': None,
                    'peak_kind': 'sustained bf16 (kernel timed inside a long step); burst peak %.1f' % tens_burst}
    else:
        ach = d['bytes'] / d['ms'] / 1e6
        roofline = {'bound': 'hbm', 'kernel': k, 'achieved': round(ach, 1), 'peak': hbm_peak, 'unit': 'GB/s',
                    'frac': round(ach / hbm_peak, 4), 'traffic': None}
    tr_path = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')        # committed ncu --set full capture, per launch
    if os.path.exists(tr_path):
        tr = json.load(open(tr_path)).get(k)
        if tr:
            roofline['traffic'] = tr['dram_read_bytes'] + tr['dram_write_bytes']
            roofline['traffic_source'] = 'profiles/ncu_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum, one launch)'
            roofline['algorithmic_bytes_per_launch'] = tr['algorithmic_bytes']
    step_flops = conv_fl / 3
    roofline.update({'share_of_step': round(d['ms'] / tot, 4), 'peak_source': src,
                     'step_algorithmic_gbs': round(sum(v['bytes'] for v in agg.values()) / tot / 1e6, 1),
                     'conv_kernels': {'share_of_step': round(conv_ms / tot, 4),
                                      'tflops': round(conv_fl / conv_ms / 1e9, 1) if conv_ms else None,
                                      'frac_of_bf16_sustained_peak': round(conv_fl / conv_ms / 1e9 / tens_sus, 4) if conv_ms else None,
                                      'frac_of_bf16_burst_peak': round(conv_fl / conv_ms / 1e9 / tens_burst, 4) if conv_ms else None},
                     'note': 'CUDA events around each C-ABI call in an eager single-stream pass after the timed region, each '
                             'iteration queued behind a device-side delay so that host launch gaps are not counted '
                             '(the step graph itself overlaps weight gradients / encoder branches on side streams); '
                             'algorithmic bytes = every operand read once + result written once at its storage dtype'})
    return dict(launches_per_step=launches_per_step, roofline=roofline, breakdown=breakdown, categories=categories,
                conv_flops_per_step=step_flops, tens_burst=tens_burst, tens_sus=tens_sus)


def measure_training(cfgname, B, S, Cc, dtype, steps, warmup, world, rank, local, sampler=None, profile=True):
    """Device-resident training throughput of one config: W untimed steps (eager warm-ups + graph capture), then exactly
    ``steps`` graph replays between barriers, CUDA events, max over ranks."""
    import torch
    import torch.distributed as dist
    from dnncancerannotator_b200.models import tf_models
    from dnncancerannotator_b200.synthetic import make_slices
    cfg = load_cfg(cfgname)
    m = getattr(tf_models, cfg['model'])(**cfg['model_options'], dtype=dtype)
    m.build((None, S, S, Cc))
    m.compile(optimizer=cfg['deploy_options']['optimizer'], loss=cfg['deploy_options']['loss'])
    if world > 1:
        m.enable_data_parallel()
    x, y = make_slices(B, S, S, Cc, seed=1234 + rank)
    xh = torch.from_numpy(x).pin_memory()
    yh = torch.from_numpy(y).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W = max(warmup, 3)
    loss0 = float(m.train_step(xh, yh))            # untimed step 1: builds the plan, uploads the batch
    plan = m.training_plan(B, S, S)
    for _ in range(W - 1):                         # untimed steps 2..W: second eager warm-up, graph capture, replays
        m._train_on_static(plan)
    barrier()
    if sampler is not None and rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        loss = m._train_on_static(plan)
    e1.record()
    barrier()
    wall = time.perf_counter() - wall0
    dev_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if (sampler is not None and rank == 0) else None
    t = torch.tensor([dev_ms], device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t) / steps
    out = dict(model=m, plan=plan, cfg=cfg, xh=xh, yh=yh, barrier=barrier, ms_per_step=ms_per_step,
               value=B * world / (ms_per_step / 1e3), loss_first=loss0, loss_last=float(loss), wall=wall, clocks=clocks,
               warmup_done=W, e0=e0, e1=e1, exchange=exchange_kind(m, world))
    if profile:        # every rank runs the pass (it contains the gradient all-reduce); rank 0 reports
        out.update(profile_step(m, plan, cfg, world))
    return out


def exchange_kind(m, world):
    """Which gradient exchange the step graph of this model contains (parallel.py)."""
    if world <= 1:
        return None
    if getattr(m, '_p2p', None) is not None:
        return ('all-reduce fused into the Adam kernel over NVLink peer memory (csrc/p2p_adam.cu: every rank reads the '
                'peers\' gradient buffers; no NCCL call in the step)')
    return 'NCCL buckets issued inside the step graph as the backward pass completes them'


def secondary_line(cfgname, B, args, world, rank, local):
    """BASELINE.json's metric has a second half ("conv tensor-pipe % of bf16 peak") that only the wide configs can
    answer: one short device-timed run of configs/<cfgname>.yaml with its own conv roofline."""
    import gc
    import torch
    r = measure_training(cfgname, B, args.size, args.channels, args.dtype, min(args.steps, 10), 3, world, rank, local)
    conv = r['roofline']['conv_kernels']
    step_tf = r['conv_flops_per_step'] / (r['ms_per_step'] * 1e9)      # conv FLOPs of a step / whole step time
    line = {'config': f'configs/{cfgname}.yaml', 'per_gpu_batch': B, 'value': r['value'], 'unit': UNIT,
            'ms_per_step': r['ms_per_step'], 'steps': min(args.steps, 10), 'warmup': 3, 'n_gpus': world,
            'launches_per_step': r['launches_per_step'], 'allreduce': r.get('exchange'),
            'conv_tensor_pipe': {'tflops': conv['tflops'], 'frac_of_burst_peak': conv['frac_of_bf16_burst_peak'],
                                 'frac_of_sustained_peak': conv['frac_of_bf16_sustained_peak'],
                                 'share_of_step': conv['share_of_step']},
            'whole_step_tflops': round(step_tf, 1), 'whole_step_frac_of_burst_peak': round(step_tf / r['tens_burst'], 4),
            'roofline': r['roofline'], 'categories': r['categories'], 'loss_first': r['loss_first'],
            'loss_last': r['loss_last']}
    m = r.pop('model')
    m.release_graphs()
    del r, m
    gc.collect()
    torch.cuda.empty_cache()
    return line


def run_forward_sweep(args):
    """BASELINE configs[4]: configs/multiresunet.yaml evaluate / inference sweep, forward only, batch 1..256.
    Per batch size: W untimed calls (fold + pack pass, eager warm-ups, graph capture), then K graph replays with the
    batch resident in HBM (CUDA events), and the same through the public call with HOST uint8 slices in and the
    probability maps read back (e2e).  With N ranks every rank runs its own replica (no communication)."""
    import gc
    import torch
    import torch.distributed as dist
    from dnncancerannotator_b200 import native as N
    from dnncancerannotator_b200 import hostmem
    from dnncancerannotator_b200.models import tf_models
    from dnncancerannotator_b200.synthetic import make_slices
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    cfg = load_cfg(args.config)
    S = args.size
    Cc = cfg['model_options'].get('n_channels', args.channels)
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
    lib = N.lib()
    rows = []
    for B in [int(b) for b in args.batches.split(',')]:
        m = getattr(tf_models, cfg['model'])(**cfg['model_options'], dtype=args.dtype)
        m.build((None, S, S, Cc))
        x8, _ = make_slices(B, S, S, Cc, seed=1234 + rank, as_uint8=True)
        xh = hostmem.pinned_like(x8, write_combined=False)
        xd = torch.from_numpy(x8).cuda().float().div_(255.0)
        for _ in range(max(args.warmup, 4)):
            m(xd)
        plan = m._plan(B, S, S)
        torch.cuda.synchronize()
        steps = args.steps if B >= 16 else args.steps * 4
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g = plan.graphs['infer']
        e0.record()
        for _ in range(steps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        # e2e: host uint8 -> device (/255 on the device) -> forward -> probabilities back on the host
        out_h = torch.empty((B, S, S, 1), dtype=torch.float32).pin_memory()
        for _ in range(2):
            out_h.copy_(m(xh), non_blocking=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(max(steps // 4, 3)):
            out_h.copy_(m(xh), non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        e2e_ms = e0.elapsed_time(e1) / max(steps // 4, 3)
        # launches and tensor-core share (one eager pass)
        m.use_cuda_graph = False
        for f in range(3):
            lib.dnnca_debug_family_count(f, 1)
        lib.dnnca_debug_launch_count(1)
        with N.Profiler() as prof:
            m(xd)
        launches = int(lib.dnnca_debug_launch_count(1))
        fam = [int(lib.dnnca_debug_family_count(f, 0)) for f in range(3)]
        agg = prof.summary()
        conv_ms = sum(v['ms'] for v in agg.values() if v['flops'])
        tot_ms = sum(v['ms'] for v in agg.values())
        # FLOP convention of SURVEY 8d: LOGICAL channel counts (the physically padded channels do not count)
        conv_fl = 0
        for op in plan.ops:
            if type(op).__name__ == '_FoldedConv':
                conv_fl += 2 * op.x.n * op.x.h * op.x.w * op.xs.c * op.cout * op.k * op.k
            elif type(op).__name__ == '_FoldedTConv':
                conv_fl += 2 * op.x.n * op.x.h * op.x.w * op.xs.c * op.cout * 4
        t = torch.tensor([ms, e2e_ms], device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])
        rows.append({'batch': B, 'ms_per_forward': round(ms, 4), 'slices_per_s': round(B * world / (ms / 1e3), 1),
                     'e2e_slices_per_s': round(B * world / (e2e_ms / 1e3), 1), 'launches': launches,
                     'conv_launches_generic_small_tcgen05': fam,
                     'whole_forward_tflops': round(conv_fl / (ms * 1e9), 1),
                     'conv_kernels_tflops': round(conv_fl / conv_ms / 1e9, 1) if conv_ms else None,
                     'conv_share_of_forward': round(conv_ms / tot_ms, 3) if tot_ms else None,
                     'gflop_per_slice_logical': round(conv_fl / B / 1e9, 2),
                     'breakdown': [{'kernel': k, 'ms': round(d['ms'], 4), 'calls': d['calls'],
                                    'gbs': round(d['bytes'] / d['ms'] / 1e6, 1) if d['ms'] else None}
                                   for k, d in sorted(agg.items(), key=lambda kv: -kv[1]['ms'])[:14]]})
        del m, plan, g, xd, xh, out_h
        gc.collect()
        torch.cuda.empty_cache()
    if rank == 0:
        best = max(rows, key=lambda r: r['slices_per_s'])
        line = {'metric': 'forward_slices_per_sec', 'value': best['slices_per_s'], 'unit': UNIT, 'n_gpus': world,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': best['ms_per_forward'], 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
                'config': {'workload': f'configs/{args.config}.yaml {cfg["model"]} forward (inference, BatchNorm folded), '
                                       f'{S}x{S}x{Cc} slices, batch sweep {args.batches}; value = best batch ({best["batch"]})',
                           'parallelism': f'replicas x{world}', 'cuda_graph': True},
                'e2e': {'value': best['e2e_slices_per_s'], 'unit': UNIT, 'h2d_bytes_per_step': best['batch'] * S * S * Cc,
                        'd2h_bytes_per_step': best['batch'] * S * S * 4},
                'gpu_launches': best['launches'] * args.steps, 'sweep': rows,
                'peaks': {'bf16_tflops_burst': peaks.get('bf16_tflops'), 'hbm_gbs': peaks.get('hbm_gbs')}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_ours(args):
    import torch
    import torch.distributed as dist
    from dnncancerannotator_b200.synthetic import make_slices

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    numa_cpus = None
    if world > 1:
        from dnncancerannotator_b200.parallel import bind_host_to_gpu
        numa_cpus = bind_host_to_gpu(local)       # before any pinned buffer exists: staging memory on the GPU's NUMA node
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world}'

    B, S, Cc = args.batch, args.size, args.channels
    sampler = ClockSampler(local)
    r = measure_training(args.config, B, S, Cc, args.dtype, args.steps, args.warmup, world, rank, local, sampler=sampler,
                         profile=not args.no_profile)
    m, plan, cfg, xh, yh, barrier = r['model'], r['plan'], r['cfg'], r['xh'], r['yh'], r['barrier']
    e0, e1 = r['e0'], r['e1']

    # ---- end-to-end through the public API with HOST buffers ----------------------------------
    # every step: H2D of that step's (x, y) from pinned host memory (Model.prefetch issues it on a copy stream so it
    # overlaps the previous step's kernels, like the reference's tf.data prefetch) + Model.train_step + a D2H read of
    # a step's loss.  The loss read lags one step so the pipeline stays full; every step's loss is read exactly once.
    def e2e_loop(xhost, yhost, steps):
        m.prefetch(xhost, yhost)
        prev = None
        for _ in range(steps):
            l = m.train_step(xhost, yhost)
            m.prefetch(xhost, yhost)
            if prev is not None:
                prev.item()
            prev = l
        prev.item()

    def timed_e2e(xhost, yhost):
        e2e_loop(xhost, yhost, 3)
        barrier()
        e0.record()
        e2e_loop(xhost, yhost, args.steps)
        e1.record()
        barrier()
        tt = torch.tensor([e0.elapsed_time(e1)], device='cuda')
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt) / args.steps

    # primary: the raw decoded slices -- uint8 image channels + uint8 label (data.py:193-206 reads PNG bytes and
    # divides by 255; here that division runs on the device) -- in pinned host memory
    from dnncancerannotator_b200 import hostmem
    x8, y8 = make_slices(B, S, S, Cc, seed=1234 + rank, as_uint8=True)
    x8h, y8h = hostmem.pinned_like(x8, write_combined=args.wc), hostmem.pinned_like(y8, write_combined=args.wc)
    e2e_ms = timed_e2e(x8h, y8h)
    h2d = int(x8h.numel() + y8h.numel())
    e2e = {'value': B * world / (e2e_ms / 1e3), 'unit': UNIT, 'ms_per_step': e2e_ms,
           'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
           'h2d_gbs_per_rank': round(h2d / e2e_ms / 1e6, 2),
           'host_buffers': hostmem.describe(x8h, write_combined=args.wc),
           'input': 'uint8 slices + uint8 labels in pinned host memory (the decoded-PNG contract of data.py:193-206; '
                    'the /255 runs on the device), H2D prefetched on a copy stream, loss read back every step'}
    # binary label masks shipped as bits (data_tail.pack_labels): 3.125 instead of 4 bytes per pixel over the host link --
    # the lever that is left where the HOST's copy rate bounds the end-to-end number (eight ranks on one memory system)
    from dnncancerannotator_b200 import data_tail
    yph = data_tail.pack_labels(y8)
    pk_ms = timed_e2e(x8h, yph)
    pk_bytes = int(x8h.numel()) + yph.nbytes
    e2e['packed_label_input'] = {'value': B * world / (pk_ms / 1e3), 'ms_per_step': pk_ms, 'h2d_bytes_per_step': pk_bytes,
                                 'h2d_gbs_per_rank': round(pk_bytes / pk_ms / 1e6, 2),
                                 'note': 'uint8 slices + bit-packed binary labels (numpy.packbits); the label buffer of the step is bit-identical'}
    if not args.no_f32_e2e:
        # same loop fed with float32 [0,1] tensors (the dtype the reference hands to Keras): 4x the PCIe bytes
        f32_ms = timed_e2e(xh, yh)
        e2e['float32_input'] = {'value': B * world / (f32_ms / 1e3), 'ms_per_step': f32_ms,
                                'h2d_bytes_per_step': int(xh.numel() * 4 + yh.numel() * 4),
                                'note': 'PCIe-bound: bytes / time = the host link rate'}

    secondary = []
    if args.secondary and args.config == 'unet':
        m.release_graphs()
        del m, plan, xh, yh, x8h, y8h
        r.pop('model'); r.pop('plan'); r.pop('xh'); r.pop('yh')
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        for name, sb in (('unet_big', 32), ('mulmo_unet', 32)):
            try:
                line2 = secondary_line(name, sb, args, world, rank, local)
            except Exception as exc:        # a failing secondary must never take the headline line down
                line2 = {'config': f'configs/{name}.yaml', 'error': f'{type(exc).__name__}: {exc}'[:300]}
            secondary.append(line2)

    if 'model' in r:
        r['model'].release_graphs()          # captured NCCL collectives must be gone before the communicator is destroyed
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        rc = cpu_reference_rate(cfg, args, steps=10, warmup=3)
        cpu = {'value': rc['value'], 'unit': UNIT, 'cores': rc['cores'], 'kind': 'port',
               'sample': rc['sample'] + '; torch-CPU restatement of the reference (TensorFlow unavailable)'}
    lps = r.get('launches_per_step')
    line = {
        'metric': METRIC, 'value': r['value'], 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'warmup_steps_run': r['warmup_done'], 'ms_per_step': r['ms_per_step'],
        'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
        'config': {'workload': f'configs/{args.config}.yaml {cfg["model"]} training step, per-GPU batch {B} of '
                               f'{S}x{S}x{Cc} slices ({args.dtype} activations, fp32 accumulate/master weights)',
                   'global_batch': B * world, 'parallelism': f'dp{world}',
                   'l2': 'working set per step (>1 GB of activations) exceeds the 126 MB L2; no flush needed',
                   'cuda_graph': True,
                   'allreduce': r.get('exchange'),
                   'host_numa_cpus': (f'{len(numa_cpus)} CPUs local to the GPU' if numa_cpus else None)},
        'e2e': e2e, 'gpu_launches': (lps or 0) * args.steps, 'launches_per_step': lps,
        'clocks': r['clocks'], 'roofline': r.get('roofline'), 'cpu_baseline': cpu, 'breakdown': r.get('breakdown'),
        'categories': r.get('categories'), 'secondary': secondary,
        'loss_first': r['loss_first'], 'loss_last': r['loss_last'], 'wall_s_timed_region': r['wall'],
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    elif a.forward:
        run_forward_sweep(a)
    else:
        run_ours(a)
