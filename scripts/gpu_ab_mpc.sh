#!/bin/bash
# A/B of MPC kernel knobs (env read once per process): CRB_MPC_PREFETCH = L2 prefetch distance in stages
OUT=gpurun_out; mkdir -p $OUT
for v in ${MPC_VARIANTS:-0 2 3 4 0}; do
  CRB_MPC_PREFETCH=$v python bench.py --steps 25 --warmup 5 --no-cpu --workload mpc 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.readline()); m=d['extra']['mpc']; print('CRB_MPC_PREFETCH=$v', 'solves/s %.3e' % m['value'], 'ms/step %.4f' % m['ms_per_step'], 'e2e %.3e' % m['e2e']['value'], 'checksum', m['solver']['checksum'])"
done | tee $OUT/ab_mpc.txt
