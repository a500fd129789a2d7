#!/bin/bash
# round 2, call g (1 GPU): codegen seam (NVRTC-instantiated marching kernels), fine-grained path, whole GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_codegen_seam.py tests/test_gpu_fine_grained.py -m gpu -q > gpurun_out/r02g_pytest_seam.log 2>&1
tail -30 gpurun_out/r02g_pytest_seam.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_pytest.log 2>&1
tail -4 gpurun_out/r02g_pytest.log
