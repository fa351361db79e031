for cfg in 20 22; do
SPAGXE_TC_CFG=$cfg timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "tensor_core or properties_at_scale or tiny" > gpurun_out/pytest_tc5_$cfg.log 2>&1; tail -1 gpurun_out/pytest_tc5_$cfg.log
done
for cfg in 10 20 21 22 23 24 25; do
  SPAGXE_TC_CFG=$cfg timeout 300 python bench.py --no-e2e --steps 10 --warmup 3 --cpu-sample-snps 0 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('cfg $cfg', round(d['value']), d['stages_ms']['score'], round(d['roofline']['frac'],4))
" | tee -a gpurun_out/tc5_variants.log
done
