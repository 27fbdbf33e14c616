#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r2l}
for r in bulk async; do
echo "== CRB_MPC_RING=$r"
CRB_MPC_RING=$r timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k mpc 2>&1 | tail -2
CRB_MPC_RING=$r CRB_MPC_VARIANT=1 timeout 300 python scripts/mpc_probe.py
done 2>&1 | tee $OUT/mpc_ring_ab_$TAG.txt
CRB_MPC_RING=bulk MPC_PROBE_CHECK=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:crb_mpc_tasks_kernel -s 3 -c 1 -f \
   -o $OUT/prof_mpc_$TAG python scripts/mpc_probe.py 65536 > $OUT/prof_mpc_$TAG.stdout 2>&1
