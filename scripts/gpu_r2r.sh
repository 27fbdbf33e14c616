#!/bin/bash
# PF three-launch iteration: parity, timing, marginal costs, one ncu --set full of the gather.
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "pf_step or pf_sharded or hinted or resample or pf_" > $OUT/r_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/r_pytest.log
for sk in 0 2 4 6; do
  CRB_PF_SKIP=$sk timeout 300 python bench.py --workload pf --no-cpu --steps 10 --warmup 3 > $OUT/r_pf_skip$sk.json 2> $OUT/r_pf_skip$sk.err
done
timeout 400 ncu --clock-control none --set full --import-source on -k regex:crb_pf_gather2 -s 3 -c 1 -f -o $OUT/r_prof_gather2 \
    python bench.py --workload pf --no-cpu --steps 4 --warmup 3 > $OUT/r_prof_gather2.stdout 2>&1
timeout 400 ncu --clock-control none --set full --import-source on -k regex:crb_pf_scan1n3 -s 3 -c 1 -f -o $OUT/r_prof_scan1n3 \
    python bench.py --workload pf --no-cpu --steps 4 --warmup 3 > $OUT/r_prof_scan1n3.stdout 2>&1
tail -3 $OUT/r_pytest.log
for sk in 0 2 4 6; do python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/r_pf_skip$sk.json") if l.startswith('{')][-1])
    print("skip $sk", "PF iter us", 1e3*d["extra"]["pf_full_iteration"]["ms_per_step"], d["extra"]["pf_full_iteration"]["timing"])
except Exception as e:
    print("skip $sk failed", e); print(open("$OUT/r_pf_skip$sk.err").read()[-800:])
PY
done
ls -la $OUT/r_prof_*.ncu-rep
