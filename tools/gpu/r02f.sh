#!/bin/bash
# round 2, call f (2 GPUs): the two-rank decomposed == single cases with the rim / interior overlap, the fine-grained C-ABI path
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fine_grained.py tests/test_slab_decomposition.py tests/test_abi.py -m gpu -q > gpurun_out/r02f_pytest_2gpu.log 2>&1
tail -15 gpurun_out/r02f_pytest_2gpu.log
