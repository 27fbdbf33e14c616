#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r2n}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k mpc 2>&1 | tail -2
{
CRB_MPC_VARIANT=1 timeout 300 python scripts/mpc_probe.py 65536 1048576
for cfg in "187 6" "211 5" "203 4"; do set -- $cfg
  CRB_MPC_VARIANT=1 CRB_MPC_SLOTS=$1 CRB_MPC_WARPS=$2 MPC_PROBE_CHECK=0 timeout 300 python scripts/mpc_probe.py 65536 1048576; done
CRB_MPC_VARIANT=0 MPC_PROBE_CHECK=0 timeout 300 python scripts/mpc_probe.py 65536 1048576
} 2>&1 | tee $OUT/mpc_probe_$TAG.txt
MPC_PROBE_CHECK=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:crb_mpc_tasks_kernel -s 3 -c 1 -f \
   -o $OUT/prof_mpc_$TAG python scripts/mpc_probe.py 65536 > $OUT/prof_mpc_$TAG.stdout 2>&1
