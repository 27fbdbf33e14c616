#!/bin/bash
# One GPU visit: parity tests, bench, ncu launch list, ncu --set full of the three hot kernels.
# Usage (through gpurun): bash scripts/gpu_round.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_$TAG.log
python bench.py --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 1200 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > $OUT/launches_$TAG.stdout 2>&1
for w in ekf pf mpc; do
  case $w in ekf) K=crb_ekf_step_kernel;; pf) K=crb_pf_predict_weight_lean_kernel;; mpc) K=crb_mpc_tasks_kernel;; esac
  $NCU --set full --import-source on -k regex:$K -s 2 -c 2 -f -o $OUT/prof_${w}_$TAG \
      python bench.py --steps 4 --warmup 3 --no-cpu --workload $w > $OUT/prof_${w}_$TAG.stdout 2>&1
done
tail -3 $OUT/pytest_gpu_$TAG.log
cat $OUT/bench_$TAG.json | head -c 6000
