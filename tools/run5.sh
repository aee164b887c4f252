#!/bin/bash
# developer run (next GPU session): A/B of the inner-loop fast group (SF_INNER_LOOP) with 4 and 3 pressure CTAs per SM against
# the committed build, then the GPU parity suite on the best candidate.  Build the variants first, in the build container:
# (or simply: bash tools/build_variants.sh)
#   python -m fluidsimulationcuda_b200.build --out build/libsf_c3.so    -DSF_PRESSURE_CTAS=3        (control: 12 warps per SM alone)
#   python -m fluidsimulationcuda_b200.build --out build/libsf_il1.so   -DSF_INNER_LOOP=1
#   python -m fluidsimulationcuda_b200.build --out build/libsf_il1c3.so -DSF_INNER_LOOP=1 -DSF_PRESSURE_CTAS=3
#   python -m fluidsimulationcuda_b200.build --out build/libsf_il2.so   -DSF_INNER_LOOP=2
#   python -m fluidsimulationcuda_b200.build --out build/libsf_il2c3.so -DSF_INNER_LOOP=2 -DSF_PRESSURE_CTAS=3
#   python -m fluidsimulationcuda_b200.build --out build/libsf_il2c3e.so -DSF_INNER_LOOP=2 -DSF_PRESSURE_CTAS=3 -DSF_EDGE_SPLIT=1
# (build/ is git-ignored but travels with the gpurun snapshot).  Static expectation (profiles/r01_sass_static.txt, "hot loop"):
# the T = 7 pressure group drops from ~605 to ~535 executed instructions (~480 with the edge split in 72 of 74 bands).
mkdir -p gpurun_out
SF_AB_T=6,7 python tools/ab_solve.py > gpurun_out/ab_r5_default.log 2>&1
for v in c3 il1 il1c3 il2 il2c3 il2c3e gg ggil2; do
  [ -f build/libsf_$v.so ] || continue
  SF_AB_T=6,7 SF_LIBRARY=$PWD/build/libsf_$v.so python tools/ab_solve.py > gpurun_out/ab_r5_$v.log 2>&1
done
cat gpurun_out/ab_r5_*.log
for v in ${SF_PARITY_VARIANTS:-il1c3}; do
  [ -f build/libsf_$v.so ] || continue
  SF_LIBRARY=$PWD/build/libsf_$v.so python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r5_$v.log 2>&1
  tail -3 gpurun_out/pytest_r5_$v.log
done
