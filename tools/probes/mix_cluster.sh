#!/bin/bash
# tuning probe: per-SNP SPAGxEmix kernel with 8 / 4 / 2 / 1 CTAs per SNP
for c in 8 4 2 1; do
  python bench.py --workload mix --steps 3 --warmup 3 --cpu-sample-snps 0 --mix-cluster $c > gpurun_out/mixc_$c.log 2>&1
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/mixc_$c.log") if l.startswith("{")][-1])
print("cluster", $c, round(d["value"]), round(d["ms_per_step"],2))
PY
done
