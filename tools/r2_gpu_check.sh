#!/bin/bash
# Round 2, GPU session 6: lane-strided advect + add_source fused into the first Jacobi launch.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp build/HEAD_COMMIT gpurun_out/s6_commit.txt 2>/dev/null || true
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/s6_pytest.log; cat gpurun_out/s6_pytest.log
SF_AB_T=7 python tools/ab_solve.py > gpurun_out/s6_ab.log 2>&1; cat gpurun_out/s6_ab.log
python tools/stage_times.py 8192 40 > gpurun_out/s6_stage_times.log 2>&1; tail -18 gpurun_out/s6_stage_times.log
python bench.py --steps 20 --warmup 5 --scaling-base 0 > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err; python -c "
import json; d=json.load(open('gpurun_out/s6_bench.json')); print('bench ms/step', d['ms_per_step'], 'launches', d['gpu_launches'], 'e2e', d['e2e'].get('ms_per_step'))"
ncu --metrics gpu__time_duration.sum --clock-control none -s 125 -c 80 --csv --log-file gpurun_out/s6_launches.csv python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/s6_ncu_launch.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none -k regex:"advect|last_project|divergence" -s 6 -c 8 --csv --log-file gpurun_out/s6_stage_ncu.csv python tools/stage_times.py 8192 40 > gpurun_out/s6_stage_ncu.log 2>&1
