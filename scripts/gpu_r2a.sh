#!/bin/bash
# Round 2, first GPU visit: parity of the resident-slot MPC kernel, A/B against the first-generation kernel,
# knob sweep, one ncu --set full capture.
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r2a}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k mpc > $OUT/pytest_mpc_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_mpc_$TAG.log
tail -5 $OUT/pytest_mpc_$TAG.log
{
CRB_MPC_VARIANT=0 timeout 300 python scripts/mpc_probe.py
CRB_MPC_VARIANT=1 timeout 300 python scripts/mpc_probe.py
for w in 4 5; do CRB_MPC_VARIANT=1 CRB_MPC_WARPS=$w MPC_PROBE_CHECK=0 timeout 300 python scripts/mpc_probe.py 65536; done
for s in 208 192 160; do CRB_MPC_VARIANT=1 CRB_MPC_SLOTS=$s MPC_PROBE_CHECK=0 timeout 300 python scripts/mpc_probe.py 65536; done
} 2>&1 | tee $OUT/mpc_probe_$TAG.txt
MPC_PROBE_CHECK=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:crb_mpc_tasks_kernel -s 3 -c 1 -f \
   -o $OUT/prof_mpc_$TAG python scripts/mpc_probe.py 65536 > $OUT/prof_mpc_$TAG.stdout 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_$TAG.log
tail -5 $OUT/pytest_gpu_$TAG.log
