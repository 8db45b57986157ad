#!/bin/bash
# tuning helper (GPU box): gpu tests, then the bench `value` leg for several register budgets
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for o in ${OCCS:-3 4 5}; do
  python bench.py --no-cpu --no-e2e --occupancy $o --steps 10 > gpurun_out/occ_$o.json
  python - <<PY
import json
d=json.load(open("gpurun_out/occ_$o.json"))
print("occ", $o, "%.4g"%d["value"], "%.3f ms"%d["ms_per_step"], "issue frac %.3f"%d["roofline"]["frac_of_nonfused_issue_peak"])
PY
done
