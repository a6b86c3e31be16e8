#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02e_bench_n2.json 2> gpurun_out/r02e_bench_n2.err; echo "bench rc=$?"
tail -5 gpurun_out/r02e_bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02e_bench_n2.json').read().strip().splitlines()[-1])
    print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['launches_per_step'])
    for s in d.get('secondary',[]):
        print(s.get('config'), s.get('value'), s.get('ms_per_step'), s.get('launches_per_step'), s.get('conv_tensor_pipe'), s.get('error'))
        for k,v in (s.get('categories') or {}).items(): print('    ',k,v)
except Exception as e: print('parse failed',e)
PY
