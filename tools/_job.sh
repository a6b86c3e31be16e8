#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "umma or tconv or conv_fprop" > gpurun_out/t_umma.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/t_umma.log
( echo "== all layers"; timeout 300 python tools/big_microbench.py --only fprop; timeout 300 python tools/big_microbench.py --only dgrad
) > gpurun_out/big_mb5.txt 2>&1
cat gpurun_out/big_mb5.txt
