#!/bin/bash
# developer run: A/B of the address-arithmetic / edge-split variants against the committed build
mkdir -p gpurun_out
for v in head edge_off64 noedge_off32; do
  SF_AB_T=7 SF_LIBRARY=$PWD/build/libsf_$v.so python tools/ab_solve.py > gpurun_out/ab_r4_$v.log 2>&1
done
SF_AB_T=5,7 python tools/ab_solve.py > gpurun_out/ab_r4_default.log 2>&1
cat gpurun_out/ab_r4_*.log
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r4.log 2>&1
tail -3 gpurun_out/pytest_r4.log
