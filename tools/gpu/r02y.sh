#!/bin/bash
# round 2, call y (1 GPU): the last binary (hb_rk_plan added, kernels unchanged): smoke, ABI, one parity sweep of the C-cases and the fold / codegen-seam tests
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02y_smoke.txt 2>&1; tail -1 gpurun_out/r02y_smoke.txt
timeout 900 python -m pytest tests/test_abi.py tests/test_rk_plan.py tests/test_gpu_codegen_seam.py tests/test_gpu_parity.py -q -k "abi or rk_plan or seam or rk_fold or C4_ or C2_ or C3_ or C5_ or march" > gpurun_out/r02y_pytest.log 2>&1; tail -3 gpurun_out/r02y_pytest.log
