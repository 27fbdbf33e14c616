#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r2j}
for t in 256 128; do
CRB_PF_GATHER_THREADS=$t timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --workload pf > $OUT/bench_pf_$TAG.json 2>$OUT/bench_pf_$TAG.err; tail -2 $OUT/bench_pf_$TAG.err
echo "gather threads $t"; python scripts/show_bench.py $OUT/bench_pf_$TAG.json | grep "PF"
done
timeout 300 nsys --version 2>/dev/null | head -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:crb_pf -s 40 -c 40 --csv --log-file $OUT/launches_pf_$TAG.csv python bench.py --steps 4 --warmup 3 --no-cpu --workload pf > /dev/null 2>&1
python scripts/launch_times.py $OUT/launches_pf_$TAG.csv
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pf" 2>&1 | tail -3
