#!/bin/bash
# the driver's SCALE command at N=8 (default bench: unet.yaml + secondary unet_big / mulmo_unet lines)
mkdir -p gpurun_out
N=${1:-8}
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02z_bench_n$N.json 2> gpurun_out/r02z_bench_n$N.err; echo "bench rc=$?"
tail -5 gpurun_out/r02z_bench_n$N.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02z_bench_n$N.json').read().strip().splitlines()[-1])
    print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'],'launches',d['launches_per_step'])
    for s in d.get('secondary',[]):
        print(s.get('config'), s.get('value'), s.get('ms_per_step'), s.get('e2e'), s.get('conv_tensor_pipe'), s.get('error'))
except Exception as e: print('parse failed',e)
PY
