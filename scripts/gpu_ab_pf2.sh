#!/bin/bash
# A/B of the software-pipelined PF kernel: variant (launch shape) x pairs per thread
OUT=gpurun_out; mkdir -p $OUT
run() { CRB_PF_VARIANT=$1 CRB_PF_PIPE_K=$2 python bench.py --steps 50 --warmup 5 --no-cpu --workload pf 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.readline()); d=d.get('extra',{}).get('pf',d); print('PF variant $1 k=$2', 'kernel GB/s %.0f frac %.3f' % (d['roofline']['achieved'], d['roofline']['frac']), 'value %.3e' % d['value'])"; }
( run 0 0; for v in 20 21 22 23; do for k in 2 3 4 6 8; do run $v $k; done; done ) | tee $OUT/ab_pf2.txt
