#!/bin/bash
# A/B of kernel variants on the GPU box: prints kernel-only roofline numbers per variant.
OUT=gpurun_out; mkdir -p $OUT
for v in 0 4 5 6 7; do
  CRB_EKF_VARIANT=$v python bench.py --steps 50 --warmup 5 --no-cpu --workload ekf 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('EKF variant $v', 'kernel GB/s %.0f frac %.3f' % (d['roofline']['achieved'], d['roofline']['frac']), 'value %.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'])"
done | tee $OUT/ab_ekf.txt
