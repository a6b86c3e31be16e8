"""``bench.py --impl reference`` on the CPU: the reference arm's JSON line carries the contract's keys (the driver computes the
GPU / CPU ratio from it), runs without a GPU and never touches the CUDA library."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra, env=None):
    e = dict(os.environ, CUDA_VISIBLE_DEVICES='')
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1',
                          '--cpu-batch', '2', '--size', '64', *extra], capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def test_reference_arm_line_schema():
    d = _run()
    assert d['impl'] == 'reference' and d['metric'] == 'train_slices_per_sec' and d['unit'] == 'slices/s'
    assert d['higher_is_better'] is True and d['n_gpus'] == 1 and d['steps'] == 1 and d['warmup'] == 1
    assert d['value'] > 0 and abs(d['ms_per_step'] - 2 / d['value'] * 1e3) < 1e-6 * d['ms_per_step'] + 1e-9
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and 'batch 2' in cb['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': 'slices/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['vs_baseline'] is None and d['data'] == 'synthetic' and d['dtype'] == 'f32'
    assert 'configs/unet.yaml' in d['config']['workload']


def test_reference_arm_under_torchrun_env_uses_all_threads_and_only_rank0_prints():
    """under torchrun (OMP_NUM_THREADS=1) the CPU arm must still use the host's threads; ranks != 0 exit 0
    without work and without a line."""
    d = _run('--gpus', '2', env=dict(OMP_NUM_THREADS='1', WORLD_SIZE='2', RANK='0', LOCAL_RANK='0'))
    assert d['cpu_baseline']['cores'] == min(os.cpu_count() or 1, d['cpu_baseline']['cores']) and d['cpu_baseline']['cores'] >= min(os.cpu_count() or 1, 2)
    e = dict(os.environ, CUDA_VISIBLE_DEVICES='', OMP_NUM_THREADS='1', WORLD_SIZE='2', RANK='1', LOCAL_RANK='1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1',
                          '--warmup', '1', '--cpu-batch', '2', '--size', '64'], capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith('{')]
