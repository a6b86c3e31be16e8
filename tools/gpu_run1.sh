#!/bin/bash
# round-2 GPU call 1: system topology, full GPU tests, real-shape parity, default bench
mkdir -p gpurun_out
{ nvidia-smi topo -m; ls /sys/devices/system/node; lscpu | head -30; nproc; free -g; } > gpurun_out/r02_sysinfo.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -15 gpurun_out/r02a_pytest.log
timeout 600 python tools/parity_real_shapes.py --out gpurun_out/r02a_parity_real_shapes.json > gpurun_out/r02a_parity.log 2>&1; echo "parity rc=$?"
tail -5 gpurun_out/r02a_parity.log | cut -c1-400
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02a_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02a_bench.json').read().strip().splitlines()[-1])
    print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'launches',d['launches_per_step'],'roof',d['roofline']['kernel'],d['roofline']['frac'])
    for s in d.get('secondary',[]): print(s.get('config'), s.get('value'), s.get('ms_per_step'), s.get('conv_tensor_pipe'), s.get('error'))
except Exception as e: print('parse failed',e)
PY
