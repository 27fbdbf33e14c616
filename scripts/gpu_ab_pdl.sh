#!/bin/bash
# A/B of programmatic dependent launch (CRB_PDL) on the two short streaming kernels
OUT=gpurun_out; mkdir -p $OUT
for pdl in 0 1 0 1; do
  CRB_PDL=$pdl python bench.py --steps 50 --warmup 5 --no-cpu --workload pf 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.readline()); p=d['extra']['pf']; print('CRB_PDL=$pdl', 'EKF frac %.3f (%.2f us/step, mode %s)' % (d['roofline']['frac'], d['ms_per_step']*1e3, d['config']['launch']), 'PF frac %.3f (%.2f us/step)' % (p['roofline']['frac'], p['ms_per_step']*1e3))"
done | tee $OUT/ab_pdl.txt
