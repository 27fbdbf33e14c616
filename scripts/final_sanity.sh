#!/bin/bash
# smoke + default bench run with wall-clock, printing the headline numbers
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
T0=$(date +%s)
python bench.py > $OUT/bench_sanity.json 2> $OUT/bench_sanity.err
echo "bench rc=$? wall=$(( $(date +%s) - T0 )) s"
python - <<PY
import json
d=json.loads([l for l in open("$OUT/bench_sanity.json") if l.startswith("{")][0])
print("EKF %.3e frac %.3f e2e %.3e"%(d["value"],d["roofline"]["frac"],d["e2e"]["value"]), "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"].get("best_effort_value"), d["cpu_baseline"]["cores"])
for k,e in d["extra"].items(): print(k, "%.3e"%e["value"], "%.3f"%e.get("roofline",{}).get("frac",0), e.get("cpu_baseline",{}).get("value"), e.get("cpu_baseline",{}).get("best_effort_value"))
PY
