#!/bin/bash
# GPU session: overlapped solves + TMA advect in the whole step -- parity, then A/B of the step time at the headline size
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "overlapped or vel_and_dens or full_size_step or run_steps" 2>&1 | tail -5 > gpurun_out/b1_pytest.log; cat gpurun_out/b1_pytest.log
timeout 600 python tools/step_ab.py 8192 40 base=19:0,16:0 tile=19:0,16:1 overlap=19:1,16:0 both=19:1,16:1 both_t6=19:1,16:6 both_t8=19:1,16:8 both_noskew=19:1,16:1,15:0 > gpurun_out/b1_step_ab.log 2>&1; cat gpurun_out/b1_step_ab.log
timeout 300 python tools/step_ab.py 4096 40 base=19:0,16:0 both=19:1,16:1 >> gpurun_out/b1_step_ab.log 2>&1; tail -2 gpurun_out/b1_step_ab.log
timeout 300 python tools/step_ab.py 1024 20 base=19:0,16:0 both=19:1,16:1 >> gpurun_out/b1_step_ab.log 2>&1; tail -2 gpurun_out/b1_step_ab.log
