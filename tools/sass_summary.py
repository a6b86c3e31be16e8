"""Opcode histogram per kernel of libdnnca.so (cuobjdump -sass): the committed proof that the conv kernels are
Blackwell-native (UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld from TMEM,
UTCBAR = tcgen05.commit -> mbarrier) -- .so / .o files are git-ignored, so the repository would otherwise hold no evidence.

  python tools/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'dnncancerannotator_b200', 'libdnnca.so')
KEYS = ['UTCHMMA', 'UTCQMMA', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'LDTM', 'STTM', 'UTCBAR', 'UTCATOMSWS', 'SYNCS', 'HMMA', 'FFMA', 'FFMA2',
        'HFMA2', 'RED', 'ATOM', 'ATOMS', 'LDG', 'STG', 'LDS', 'STS', 'SHFL', 'BAR', 'ELECT', 'ACQBULK', 'CCTL']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r'\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m and cur:
            op = m.group(1)
            kernels[cur][op] += 1
            kernels[cur]['__total__'] += 1
    try:
        demangled = subprocess.run(['c++filt'], input='\n'.join(kernels), capture_output=True, text=True).stdout.splitlines()
    except OSError:
        demangled = list(kernels)
    arch = subprocess.run(['cuobjdump', '-lelf', LIB], capture_output=True, text=True).stdout
    print('# libdnnca.so SASS opcode histogram per kernel (tools/sass_summary.py)')
    print('# ELF images:', ', '.join(sorted(set(re.findall(r'sm_\w+', arch)))))
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    print('# whole library:', ' '.join(f'{k}={tot[k]}' for k in KEYS if tot[k]), f'instructions={tot["__total__"]}')
    print('# kernels:', len(kernels), ' with UTCHMMA (tcgen05.mma):', sum(1 for c in kernels.values() if c['UTCHMMA']),
          ' with UTMALDG (TMA load):', sum(1 for c in kernels.values() if c['UTMALDG']))
    print()
    for (name, c), dm in zip(kernels.items(), demangled):
        short = re.sub(r'\(.*', '', dm)
        short = short.replace('dnnca::', '').replace('void ', '')
        cols = ' '.join(f'{k}={c[k]}' for k in KEYS if c[k])
        print(f'{short[:110]:110s} n={c["__total__"]:6d}  {cols}')


if __name__ == '__main__':
    sys.exit(main())
