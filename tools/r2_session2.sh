#!/bin/bash
# Round 2, GPU session 2: parity suite + bench on the build that adopted the A/B winners.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp build/HEAD_COMMIT gpurun_out/s2_commit.txt 2>/dev/null || true
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > gpurun_out/s2_pytest.log; cat gpurun_out/s2_pytest.log
SF_AB_T=7 python tools/ab_solve.py > gpurun_out/s2_ab.log 2>&1; cat gpurun_out/s2_ab.log
python tools/stage_times.py 8192 40 > gpurun_out/s2_stage_times.log 2>&1; tail -18 gpurun_out/s2_stage_times.log
