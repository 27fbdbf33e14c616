#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "pf_step or pf_sharded or resample or pf_" > $OUT/u_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/u_pytest.log
for sk in 0 4; do
  CRB_PF_SKIP=$sk timeout 300 python bench.py --workload pf --no-cpu --steps 10 --warmup 3 > $OUT/u_pf_skip$sk.json 2> $OUT/u_pf_skip$sk.err
done
tail -3 $OUT/u_pytest.log
for f in u_pf_skip0 u_pf_skip4; do python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/$f.json") if l.startswith('{')][-1])
    print("$f", "PF iter us", 1e3*d["extra"]["pf_full_iteration"]["ms_per_step"])
except Exception as e:
    print("$f failed", e); print(open("$OUT/$f.err").read()[-800:])
PY
done
