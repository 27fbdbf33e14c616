#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r2g}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pf or ekf or mpc_host" > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_$TAG.log; tail -4 $OUT/pytest_$TAG.log
for t in 1 0; do
  CRB_PF_TRIG=$t timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --workload pf > $OUT/bench_pf_trig$t.json 2>/dev/null
  python -c "
import json; d=json.loads(open('$OUT/bench_pf_trig$t.json').readline()); p=d['extra']['pf']; print('CRB_PF_TRIG=$t', 'PF %.4e particles/s frac %.3f' % (p['value'], p['roofline']['frac']), 'EKF frac %.3f' % d['roofline']['frac'], 'PF iter ms', d['extra']['pf_full_iteration']['ms_per_step'])"
done 2>&1 | tee $OUT/pf_trig_ab_$TAG.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --workload mpc > $OUT/bench_mpc_$TAG.json 2>$OUT/bench_mpc_$TAG.err; python scripts/show_bench.py $OUT/bench_mpc_$TAG.json | grep MPC
{
for cfg in "187 6" "195 5" "211 5"; do set -- $cfg
  CRB_MPC_VARIANT=1 CRB_MPC_SLOTS=$1 CRB_MPC_WARPS=$2 MPC_PROBE_CHECK=0 timeout 300 python scripts/mpc_probe.py 65536 1048576; done
} 2>&1 | tee $OUT/mpc_probe_$TAG.txt
MPC_PROBE_CHECK=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:crb_mpc_tasks_kernel -s 3 -c 1 -f \
   -o $OUT/prof_mpc_$TAG python scripts/mpc_probe.py 65536 > $OUT/prof_mpc_$TAG.stdout 2>&1
