"""In-tree nvcc build of libdnnca.so for sm_100a.

``python -m dnncancerannotator_b200.build`` (or ``__graft_entry__.build()``)
compiles every ``csrc/*.cu`` translation unit in parallel with
``-gencode arch=compute_100a,code=sm_100a -lineinfo`` and links them into
``dnncancerannotator_b200/libdnnca.so``.  The .so stays in-tree (git-ignored) so
it travels to the GPU box with the repo snapshot; objects are cached under
``csrc/_build`` and rebuilt when a source or header is newer.
"""
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(CSRC, '_build')
LIB = os.path.join(HERE, 'libdnnca.so')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '-I' + os.path.join(ROOT, 'include'),
              '-I' + CSRC, '-DDNNCA_BUILD']

# (source, object tag, extra defines)
UNITS = [
    ('api.cu', 'api', []),
    ('elementwise.cu', 'elementwise', []),
    ('conv_generic.cu', 'conv_generic', []),
    ('head_loss.cu', 'head_loss', []),
    ('optim.cu', 'optim', []),
    ('metrics.cu', 'metrics', []),
    ('region_metrics.cu', 'region_metrics', []),
    ('tconv_small.cu', 'tconv_small', []),
    ('input_tail.cu', 'input_tail', []),
    ('tps_warp.cu', 'tps_warp', []),
    ('bn_fold.cu', 'bn_fold', []),
    ('p2p_adam.cu', 'p2p_adam', []),
    ('nccl_wrap.cu', 'nccl_wrap', []),
    ('multires_train.cu', 'multires_train', []),
]
for dt in (0, 1):
    for kind in (0, 1, 2):
        UNITS.append(('conv_small.cu', f'conv_small_d{dt}k{kind}', [f'-DSMALL_DT={dt}', f'-DSMALL_KIND={kind}']))
UNITS.append(('conv_umma.cu', 'conv_umma', []))
UNITS.append(('conv_umma2.cu', 'conv_umma2', []))
UNITS.append(('conv_umma3.cu', 'conv_umma3', []))
UNITS.append(('conv_row_umma.cu', 'conv_row_umma', []))
UNITS.append(('pool_vec.cu', 'pool_vec', []))


def nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found: libdnnca.so cannot be built (there is no CPU fallback)')
    return exe


def _newest_header():
    t = 0.0
    for d in (CSRC, os.path.join(ROOT, 'include')):
        for f in os.listdir(d):
            if f.endswith(('.cuh', '.h')):
                t = max(t, os.path.getmtime(os.path.join(d, f)))
    return max(t, os.path.getmtime(os.path.abspath(__file__)))


def _compile(unit, hdr_time, verbose):
    src, tag, defs = unit
    srcp, objp = os.path.join(CSRC, src), os.path.join(OBJ, tag + '.o')
    if os.path.exists(objp) and os.path.getmtime(objp) >= max(os.path.getmtime(srcp), hdr_time):
        return objp, False
    cmd = [nvcc()] + NVCC_FLAGS + defs + ['-c', srcp, '-o', objp]
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'nvcc failed for {src} ({tag}):\n{r.stdout}\n{r.stderr}')
    if verbose:
        sys.stderr.write(r.stderr)
    return objp, True


def build(force=False, verbose=False, jobs=None):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    hdr_time = _newest_header()
    jobs = jobs or min(len(UNITS), os.cpu_count() or 4)
    with cf.ThreadPoolExecutor(jobs) as ex:
        results = list(ex.map(lambda u: _compile(u, hdr_time, verbose), UNITS))
    objs = [o for o, _ in results]
    if any(c for _, c in results) or not os.path.exists(LIB):
        cmd = [nvcc(), '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB] + objs + ['-lcudart', '-ldl']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
