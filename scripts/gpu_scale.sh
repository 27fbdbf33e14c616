#!/bin/bash
# N-GPU visit: multi-GPU tests, then bench.py at N = all GPUs of the box through torchrun
OUT=gpurun_out; mkdir -p $OUT
N=${1:-2}; TAG=${2:-r2}
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_multi_gpu_cpp.py tests/test_gpu_parity.py -m gpu -x -q -k "two_gpus or sharded" 2>&1 | tail -3
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n${N}_$TAG.json 2> $OUT/bench_n${N}_$TAG.err; echo "bench n$N rc=$?"; tail -2 $OUT/bench_n${N}_$TAG.err
python scripts/show_bench.py $OUT/bench_n${N}_$TAG.json
