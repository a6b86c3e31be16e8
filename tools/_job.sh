#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/t_gpu.log
timeout 300 python bench.py --config unet_big --batch 32 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/h_big32.json 2> gpurun_out/h_big32.err; echo "bench big32 rc=$?"
timeout 300 python bench.py --config unet_big --batch 16 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/h_big16.json 2> gpurun_out/h_big16.err; echo "bench big16 rc=$?"
timeout 300 python bench.py --config mulmo_unet --batch 32 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/h_mulmo.json 2> gpurun_out/h_mulmo.err; echo "bench mulmo rc=$?"
timeout 600 python bench.py > gpurun_out/h_unet.json 2> gpurun_out/h_unet.err; echo "bench unet rc=$?"
python - <<'PY'
import json
for f in ['h_big32','h_big16','h_mulmo','h_unet']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],3), d['roofline'].get('conv_kernels'), 'e2e', d.get('e2e',{}).get('value'))
        print({k:(v['ms'],v['tflops']) for k,v in d['categories'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
