#!/bin/bash
# PF three-launch iteration: batched prologue loads, conflict-free moments, carve-out preference (A/B) + ncu of both kernels.
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "pf_step or pf_sharded or resample or pf_" > $OUT/t_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/t_pytest.log
for sk in 0 4 6; do
  CRB_PF_SKIP=$sk timeout 300 python bench.py --workload pf --no-cpu --steps 10 --warmup 3 > $OUT/t_pf_skip$sk.json 2> $OUT/t_pf_skip$sk.err
done
CRB_PF_CARVEOUT=0 timeout 300 python bench.py --workload pf --no-cpu --steps 10 --warmup 3 > $OUT/t_pf_nocarve.json 2> $OUT/t_pf_nocarve.err
CRB_PDL=0 timeout 300 python bench.py --workload pf --no-cpu --steps 10 --warmup 3 > $OUT/t_pf_nopdl.json 2> $OUT/t_pf_nopdl.err
for K in gather2 scan1n3; do
timeout 300 ncu --clock-control none --set full --import-source on -k regex:crb_pf_$K -s 3 -c 1 -f -o $OUT/t_prof_$K \
    python bench.py --workload pf --no-cpu --steps 4 --warmup 3 > $OUT/t_prof_$K.stdout 2>&1
done
tail -3 $OUT/t_pytest.log
for f in t_pf_skip0 t_pf_skip4 t_pf_skip6 t_pf_nocarve t_pf_nopdl; do python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/$f.json") if l.startswith('{')][-1])
    print("$f", "PF iter us", 1e3*d["extra"]["pf_full_iteration"]["ms_per_step"])
except Exception as e:
    print("$f failed", e); print(open("$OUT/$f.err").read()[-800:])
PY
done
