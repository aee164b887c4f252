#!/bin/bash
# A/B of the packed-f32x2 Jacobi arithmetic against the scalar build (developer tool, one GPU)
mkdir -p gpurun_out
tools/micro/f32x2bench > gpurun_out/f32x2bench.log 2>&1
SF_LIBRARY=$PWD/build/libsf_base.so python tools/ab_solve.py > gpurun_out/ab_base.log 2>&1
python tools/ab_solve.py > gpurun_out/ab_packed.log 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_packed.log 2>&1
tail -3 gpurun_out/pytest_packed.log
cat gpurun_out/f32x2bench.log gpurun_out/ab_base.log gpurun_out/ab_packed.log
