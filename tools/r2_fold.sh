#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "overlapped or full_size_step" 2>&1 | tail -8 > gpurun_out/d1_pytest.log; cat gpurun_out/d1_pytest.log
for i in 1 2; do timeout 300 python tools/step_ab.py 8192 40 folded= two_kernels=20:0 folded_again= two_again=20:0; done > gpurun_out/d1_fold_ab.log 2>&1; cat gpurun_out/d1_fold_ab.log
timeout 300 python tools/step_ab.py 4096 40 folded= two_kernels=20:0 >> gpurun_out/d1_fold_ab.log 2>&1; tail -2 gpurun_out/d1_fold_ab.log
