#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "overlapped" 2>&1 | tail -5 > gpurun_out/b3_pytest.log; cat gpurun_out/b3_pytest.log
for rep in 1 2; do
echo "== density solve forked after the velocity solves"
timeout 300 python tools/step_ab.py 8192 40 off=19:0 skew=19:1 noskew=19:1,15:0
echo "== density solve forked at the start"
SF_DEV_DENS_EARLY=1 timeout 300 python tools/step_ab.py 8192 40 skew=19:1 noskew=19:1,15:0
done > gpurun_out/b3_overlap_order.log 2>&1; cat gpurun_out/b3_overlap_order.log
