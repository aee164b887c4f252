#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "overlapped" 2>&1 | tail -5 > gpurun_out/b2_pytest.log; cat gpurun_out/b2_pytest.log
for p in hlh lll hhh hll lhh llh hhl lhl; do   # lane0 (v solve), lane1 (density solve), main (u solve, projections, advects)
  echo "== priorities lane0/lane1/main = $p"
  SF_DEV_LANE_PRIO=$p timeout 300 python tools/step_ab.py 8192 40 skew=16:6 noskew=16:6,15:0
done > gpurun_out/b2_prio.log 2>&1; cat gpurun_out/b2_prio.log
timeout 600 python tools/step_ab.py 8192 40 s0=16:6,15:0 s110100=16:6,15:110100 s120100=16:6,15:120100 s115105=16:6,15:115105 s105100=16:6,15:105100 s125110=16:6,15:125110 > gpurun_out/b2_skew.log 2>&1; cat gpurun_out/b2_skew.log
