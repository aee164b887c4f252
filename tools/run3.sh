#!/bin/bash
# developer run: parity of the group-vote build, pressure-plan A/B, ncu counters of the Jacobi kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r3.log 2>&1
tail -3 gpurun_out/pytest_r3.log
SF_AB_T=7,8 python tools/ab_solve.py > gpurun_out/ab_r3.log 2>&1
SF_AB_T=7 SF_PRESSURE_PLAN=0 python tools/ab_solve.py > gpurun_out/ab_r3_evenplan.log 2>&1
cat gpurun_out/ab_r3.log gpurun_out/ab_r3_evenplan.log
for cfg in "strict 7 8192 1" "strict 7 8192 0" "pressure 8 8192 0"; do
  tag=r3_$(echo $cfg | tr ' ' '_')
  ncu --set full --clock-control none --import-source on -k regex:jacobi_stream -s 4 -c 1 -o gpurun_out/$tag -f python tools/prof_solve.py $cfg > gpurun_out/$tag.log 2>&1
  ncu -i gpurun_out/$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_all.csv 2>/dev/null
  rm -f gpurun_out/$tag.ncu-rep
done
ls -la gpurun_out/r3_*
