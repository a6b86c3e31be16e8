#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/t_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/i_unet.json 2> gpurun_out/i_unet.err; echo "bench unet rc=$?"
python - <<'PY'
import json
for f in ['i_unet']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', d.get('e2e',{}).get('value'))
        print({k:(v['ms'],v['gbs']) for k,v in d['categories'].items()})
        for k in d['breakdown'][:40]:
            if '@64' in k['kernel'] or '@32' in k['kernel']: print(k['kernel'], k['ms'], k['gbs'])
    except Exception as e: print(f, 'ERR', e)
PY
