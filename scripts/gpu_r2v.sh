#!/bin/bash
# ncu metrics of one HINTED launch of the MPC task kernel (config 4): lanes per instruction, traffic, issue utilisation.
OUT=gpurun_out; mkdir -p $OUT
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,sm__cycles_elapsed.avg
timeout 280 ncu --clock-control none --metrics $M -k regex:crb_mpc_tasks_kernel -s 96 -c 2 --csv --log-file $OUT/v_mpc_hinted.csv \
    python bench.py --workload mpc --no-cpu --steps 20 --warmup 5 > $OUT/v_mpc_hinted.stdout 2>&1
grep -v "^==" $OUT/v_mpc_hinted.csv | cut -d, -f5,12- | head -30
