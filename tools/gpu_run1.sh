#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
tail -25 gpurun_out/r02d_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02d_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02d_bench.json').read().strip().splitlines()[-1])
    print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['launches_per_step'],'roof',d['roofline']['kernel'],d['roofline']['frac'])
    for s in d.get('secondary',[]):
        print(s.get('config'), s.get('value'), s.get('ms_per_step'), s.get('launches_per_step'), s.get('conv_tensor_pipe'), s.get('error'))
        for k,v in (s.get('categories') or {}).items(): print('    ',k,v)
except Exception as e: print('parse failed',e)
PY
timeout 300 python bench.py --config multiresunet --forward --steps 10 --warmup 4 --batches 32 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        for r in d['sweep']:
            b=r.pop('breakdown'); print(r)
            for k in b[:10]: print('   ',k)
"
