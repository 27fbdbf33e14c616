#!/bin/bash
# full GPU test-suite, default bench line (N=1), MPC probe
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r2f}
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_$TAG.log
tail -6 $OUT/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
tail -3 $OUT/bench_$TAG.err
python scripts/show_bench.py $OUT/bench_$TAG.json
{
CRB_MPC_VARIANT=1 MPC_PROBE_CHECK=0 timeout 300 python scripts/mpc_probe.py
CRB_MPC_VARIANT=1 CRB_MPC_WARPS=4 MPC_PROBE_CHECK=0 timeout 300 python scripts/mpc_probe.py
} 2>&1 | tee $OUT/mpc_probe_$TAG.txt
