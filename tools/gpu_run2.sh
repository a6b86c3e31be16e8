#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_boundary.py -m gpu -q -k "multires or save_and_load or input_tail" 2>&1 | tail -30
timeout 900 python bench.py --config multiresunet --forward --steps 10 --warmup 4 --batches 32 > gpurun_out/r02b_multires_sweep.json 2> gpurun_out/r02b_multires_sweep.err; echo "sweep rc=$?"
tail -5 gpurun_out/r02b_multires_sweep.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02b_multires_sweep.json').read().strip().splitlines()[-1])
    for r in d['sweep']:
        b=r.pop('breakdown'); print(r)
        for k in b: print('   ',k)
except Exception as e: print('parse failed', e)
PY
