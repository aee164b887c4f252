#!/bin/bash
# Evidence on one B200 for the current build: GPU tests, the bench line, the ncu launch list, full captures of the Jacobi and
# advect kernels, the depth sweep and the staging A/B.  Outputs in gpurun_out/final_*; copy what is to be judged into profiles/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp build/HEAD_COMMIT gpurun_out/final_commit.txt 2>/dev/null || true
if [ -z "$ONLY_NCU" ]; then
[ -n "$SKIP_PYTEST" ] || { timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/final_pytest.log; cat gpurun_out/final_pytest.log; }
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 300 gpurun_out/final_bench.json; echo
python __graft_entry__.py smoke 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -s 110 -c 74 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/final_ncu_launch.log 2>&1
fi
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,launch__waves_per_multiprocessor,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__cycles_active.avg,sm__cycles_elapsed.avg,sm__cycles_elapsed.avg.per_second,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__throughput.avg.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio
# one full step (graphs off: every kernel of the step is a separate launch), full set on the kernels that matter
# (ncu prints template arguments as "(int)7, (int)0, (int)0"; the skip counts put every capture into the 3rd / 4th step of
# tools/prof_step.py: in step 0 the velocity right-hand side is dt * source with 1 % exact zeros, which is not what a timed step sees)
for spec in "20:jacobi_stream_kernel<.int.7, .int.0, .int.0>" "25:jacobi_stream_kernel<.int.8, .int.1, .int.0>" "8:jacobi_stream_kernel<.int.7, .int.0, .int.3>" "5:jacobi_stream_kernel<.int.7, .int.0, .int.6>" "2:advect_tile_kernel<.int.2" "2:advect_tile_kernel<.int.1" "1:last_project4" "1:divergence4" "2:init4"; do
  skip=${spec%%:*}; pat=${spec#*:}
  tag=$(echo "$pat" | tr -c 'A-Za-z0-9' '_' | sed 's/__*/_/g; s/_$//')
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$pat" -s $skip -c 1 -o gpurun_out/final_prof_$tag -f python tools/prof_step.py > gpurun_out/final_prof_$tag.log 2>&1
  ncu -i gpurun_out/final_prof_$tag.ncu-rep --page raw --csv --metrics $M > gpurun_out/final_prof_$tag.csv 2>/dev/null
  [ "$tag" = "jacobi_stream_kernel_int_7_int_0_int_0" ] || rm -f gpurun_out/final_prof_$tag.ncu-rep     # gpurun brings back at most 64 MiB
done
[ -z "$ONLY_NCU" ] || exit 0
# BASELINE config 3: lin_solve time and effective bandwidth against the temporal-blocking depth
python tools/t_sweep.py 8192 40 1,2,3,4,5,6,7,8 > gpurun_out/final_t_sweep.log 2>&1; cat gpurun_out/final_t_sweep.log
# staging A/B: per-lane cp.async (default) against one bulk copy per row piece (cp.async.bulk, SF_OPT_STAGING = 1)
for st in 0 1; do python tools/solve_sweep.py $st; done > gpurun_out/final_staging_ab.log 2>&1; cat gpurun_out/final_staging_ab.log
python tools/stage_times.py 8192 40 > gpurun_out/final_stage_times.log 2>&1; tail -18 gpurun_out/final_stage_times.log
# the step against the two round-2 options that changed its schedule: overlapped solves (19), TMA-staged advect (16)
python tools/step_ab.py 8192 40 sequential_gather=19:0,16:0 tile_only=19:0,16:1 overlap_only=19:1,16:0 default=19:1,16:1 > gpurun_out/final_step_ab.log 2>&1; cat gpurun_out/final_step_ab.log
python tools/advect_ab.py 8192 40 0,1,15,8,6 > gpurun_out/final_advect_ab.log 2>&1; cat gpurun_out/final_advect_ab.log
ls -la gpurun_out/final_*
