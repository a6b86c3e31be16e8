#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "umma or tconv or conv_fprop" > gpurun_out/t_umma.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/t_umma.log
( echo "== all layers"; timeout 300 python tools/big_microbench.py --only fprop; timeout 300 python tools/big_microbench.py --only dgrad
echo "== tconv"; timeout 120 python tools/tconv_microbench.py
) > gpurun_out/big_mb4.txt 2>&1
cat gpurun_out/big_mb4.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_umma_halo -s 2 -c 1 -o gpurun_out/halo_L1b -f python tools/big_microbench.py --only fprop --layers 1 --reps 1 > gpurun_out/ncu1.log 2>&1; echo "rc=$?"
