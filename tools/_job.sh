#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "row" > gpurun_out/t_row.log 2>&1; echo "row tests rc=$?"; tail -3 gpurun_out/t_row.log
for m in 0 1; do echo "== tma_store=$m"; if [ $m = 1 ]; then export DNNCA_ROW_TMA_STORE=1; fi; timeout 200 python tools/conv_microbench.py --layers 0,3,5,6,10 --only fprop 2>&1 | tail -6;  timeout 200 python tools/conv_microbench.py --layers 0,3,5,6,10 --only dgrad 2>&1 | tail -6; timeout 300 python bench.py --no-cpu-baseline --no-profile 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"; done > gpurun_out/ds.txt 2>&1
cat gpurun_out/ds.txt
