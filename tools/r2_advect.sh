#!/bin/bash
# GPU session: the TMA-staged advect kernel -- parity, then A/B at the headline size (two register budgets)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_advect_tile_gpu.py -x -q 2>&1 | tail -15 > gpurun_out/a1_pytest.log; cat gpurun_out/a1_pytest.log
timeout 300 python tools/advect_ab.py 8192 40 0,15,14,16,18,8,6 > gpurun_out/a1_ab_default.log 2>&1; cat gpurun_out/a1_ab_default.log
for v in build/at_minb2.so; do
  [ -f $v ] && { echo "== $v"; SF_LIBRARY=$v timeout 300 python tools/advect_ab.py 8192 40 0,8,6 2>&1 | tee gpurun_out/a1_ab_$(basename $v .so).log; }
done
timeout 300 python tools/advect_ab.py 4096 40 0,15,14,8,6 > gpurun_out/a1_ab_4096.log 2>&1; cat gpurun_out/a1_ab_4096.log
