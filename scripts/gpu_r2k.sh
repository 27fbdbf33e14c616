#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r2k}
for f in 0 1 0 1; do
CRB_PF_FUSE=$f timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --workload pf > $OUT/bench_pf_$TAG.json 2>$OUT/bench_pf_$TAG.err; tail -2 $OUT/bench_pf_$TAG.err
echo "CRB_PF_FUSE=$f"; python scripts/show_bench.py $OUT/bench_pf_$TAG.json | grep "PF"
done 2>&1 | tee $OUT/pf_fuse_ab_$TAG.txt
