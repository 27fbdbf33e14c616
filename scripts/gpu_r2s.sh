#!/bin/bash
# PF three-launch iteration after the ncu pass (async tile in normalise+scan, cheaper gather): parity + timing.
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "pf_step or pf_sharded or resample or pf_" > $OUT/s_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/s_pytest.log
for sk in 0 4 6; do
  CRB_PF_SKIP=$sk timeout 300 python bench.py --workload pf --no-cpu --steps 10 --warmup 3 > $OUT/s_pf_skip$sk.json 2> $OUT/s_pf_skip$sk.err
done
tail -3 $OUT/s_pytest.log
for sk in 0 4 6; do python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/s_pf_skip$sk.json") if l.startswith('{')][-1])
    print("skip $sk", "PF iter us", 1e3*d["extra"]["pf_full_iteration"]["ms_per_step"], "pf frac", d["roofline"]["pf"]["frac"])
except Exception as e:
    print("skip $sk failed", e); print(open("$OUT/s_pf_skip$sk.err").read()[-800:])
PY
done
