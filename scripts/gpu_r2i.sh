#!/bin/bash
# 2-GPU visit: full GPU test-suite (incl. the multi-GPU tests), PF timing, bench at N=2 through torchrun
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r2i}
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_$TAG.log; tail -8 $OUT/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --workload pf > $OUT/bench_pf_$TAG.json 2>$OUT/bench_pf_$TAG.err; tail -2 $OUT/bench_pf_$TAG.err
python scripts/show_bench.py $OUT/bench_pf_$TAG.json | grep "PF"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2_$TAG.json 2> $OUT/bench_n2_$TAG.err; echo "bench n2 rc=$?"; tail -3 $OUT/bench_n2_$TAG.err
python scripts/show_bench.py $OUT/bench_n2_$TAG.json
