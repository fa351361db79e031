rm -f gpurun_out/tc5_variants.log
for cfg in 10 21 22 23 24 25; do
  SPAGXE_TC_CFG=$cfg timeout 120 python bench.py --no-e2e --steps 10 --warmup 3 --cpu-sample-snps 0 2>gpurun_out/err_$cfg.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('cfg $cfg', round(d['value']), d['stages_ms']['score'], round(d['roofline']['frac'],4))
" | tee -a gpurun_out/tc5_variants.log
  tail -2 gpurun_out/err_$cfg.log | cut -c1-300
done
for cfg in 21 23; do
SPAGXE_TC_CFG=$cfg timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "tensor_core or properties_at_scale or tiny" > gpurun_out/pytest_tc5_$cfg.log 2>&1; tail -1 gpurun_out/pytest_tc5_$cfg.log
done
SPAGXE_TC_CFG=24 timeout 240 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -k "tensor_core" > gpurun_out/sanitizer_24.log 2>&1; grep -m5 -A12 "=========" gpurun_out/sanitizer_24.log | head -40
