#!/bin/bash
# Round 2, GPU session 1: re-ground every number on HEAD (tools/final_1gpu.sh), A/B the build-option variants
# (tools/run5.sh, parity suite on every variant that exists), per-stage times, ncu of the stage kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp build/HEAD_COMMIT gpurun_out/s1_commit.txt 2>/dev/null || true
bash tools/final_1gpu.sh > gpurun_out/s1_final.log 2>&1
SF_PARITY_VARIANTS="il2c3 gg ggil2 il1" bash tools/run5.sh > gpurun_out/s1_run5.log 2>&1
python tools/stage_times.py 8192 40 > gpurun_out/s1_stage_times.log 2>&1
# stage kernels of one step under ncu (advect, last_project, divergence, add_source): DRAM bytes, L1 sectors per request
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none -k regex:"advect|last_project|divergence|add_source" -s 12 -c 12 --csv --log-file gpurun_out/s1_stage_ncu.csv python tools/stage_times.py 8192 40 > gpurun_out/s1_stage_ncu.log 2>&1
tail -5 gpurun_out/s1_final.log; tail -30 gpurun_out/s1_run5.log; tail -20 gpurun_out/s1_stage_times.log
