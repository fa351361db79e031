#!/bin/bash
# tuning probe: score-kernel geometries of the probe build (groups x A buffers per group); prints ms of the score stage
for c in 1 31 41 61 22 23 10; do
  timeout 120 python bench.py --steps 10 --warmup 3 --no-e2e --cpu-sample-snps 0 --probe-cfg $c > gpurun_out/scorev_$c.log 2>&1
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/scorev_$c.log") if l.startswith("{")][-1])
    print("cfg", $c, "score ms", round(d["stages_ms"]["score"],4), "frac", round(d["roofline"]["frac"],4), "n_spa", d["routing"]["n_spa"], "newton", d["routing"]["newton_iters"])
except Exception as ex:
    print("cfg", $c, "FAILED", ex)
PY
done
