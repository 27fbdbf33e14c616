#!/bin/bash
# MPC hinted order through a counting-sort permutation + launch list of the PF iteration kernels.
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "hinted or full_size_mpc" > $OUT/p_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/p_pytest.log
timeout 400 python bench.py --workload mpc --no-cpu --steps 20 --warmup 5 > $OUT/p_mpc.json 2> $OUT/p_mpc.err
timeout 400 ncu --clock-control none --metrics gpu__time_duration.sum -k regex:crb_pf -c 120 --csv --log-file $OUT/p_pf_launches.csv \
    python bench.py --workload pf --no-cpu --steps 3 --warmup 2 > $OUT/p_pf_launches.stdout 2>&1
tail -4 $OUT/p_pytest.log
python scripts/show_bench.py $OUT/p_mpc.json 2>&1 | grep "MPC" || tail -20 $OUT/p_mpc.err
python scripts/launch_times.py $OUT/p_pf_launches.csv 2>&1 | head -20
