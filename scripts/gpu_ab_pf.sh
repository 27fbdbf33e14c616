#!/bin/bash
# A/B of the PF predict+weight kernels on the GPU box (CRB_PF_VARIANT, see crb_pf.cu pf_launch).
OUT=gpurun_out; mkdir -p $OUT
for v in ${PF_VARIANTS:-0 15 3 2 1}; do
  CRB_PF_VARIANT=$v python bench.py --steps 50 --warmup 5 --no-cpu --workload pf 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.readline()); d=d.get('extra',{}).get('pf',d); print('PF variant $v', 'kernel GB/s %.0f frac %.3f' % (d['roofline']['achieved'], d['roofline']['frac']), 'value %.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'])"
done | tee $OUT/ab_pf.txt
