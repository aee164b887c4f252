#!/bin/bash
# developer run: parity of the packed build, stealing scope A/B, ncu captures of the packed Jacobi kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_packed.log 2>&1
tail -3 gpurun_out/pytest_packed.log
SF_AB_T=7 python tools/ab_solve.py > gpurun_out/ab_packed.log 2>&1
SF_AB_T=7 SF_STEAL_SCOPE=1 python tools/ab_solve.py > gpurun_out/ab_packed_stealall.log 2>&1
cat gpurun_out/ab_packed.log gpurun_out/ab_packed_stealall.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,launch__waves_per_multiprocessor,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__cycles_active.avg,sm__cycles_elapsed.avg,sm__cycles_elapsed.avg.per_second,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_drain_per_issue_active.ratio,smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio,smsp__average_warps_issue_stalled_selected_per_issue_active.ratio,sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active
for cfg in "strict 7 8192 1" "pressure 7 8192 0"; do
  tag=r2_$(echo $cfg | tr ' ' '_')
  ncu --set full --clock-control none --import-source on -k regex:jacobi_stream -s 4 -c 1 -o gpurun_out/$tag -f python tools/prof_solve.py $cfg > gpurun_out/$tag.log 2>&1
  ncu -i gpurun_out/$tag.ncu-rep --page raw --csv --metrics $M > gpurun_out/$tag.csv 2>/dev/null
  ncu -i gpurun_out/$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_all.csv 2>/dev/null
  rm -f gpurun_out/$tag.ncu-rep
done
ls -la gpurun_out/r2_*
