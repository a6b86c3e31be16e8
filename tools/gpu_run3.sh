#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_dp_nccl.py -m gpu -x -q > gpurun_out/r02c_dp_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/r02c_dp_pytest.log
tail -n 6 gpurun_out/dp_test_*.log
timeout 500 python -m pytest tests/test_gpu_bnfold.py tests/test_gpu_models.py -m gpu -x -q 2>&1 | tail -40
