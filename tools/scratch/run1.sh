set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "tensor_core or properties_at_scale or tiny" > gpurun_out/pytest_tc4.log 2>&1; tail -3 gpurun_out/pytest_tc4.log
for cfg in 10 13 14 15; do
  SPAGXE_TC_CFG=$cfg timeout 300 python bench.py --no-e2e --steps 10 --warmup 3 --cpu-sample-snps 0 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('cfg $cfg', d['value'], d['stages_ms'], d['roofline']['frac'])
"
done
