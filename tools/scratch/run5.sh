timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest.log 2>&1; tail -2 gpurun_out/pytest.log
timeout 200 python bench.py --no-e2e --steps 10 --warmup 3 --cpu-sample-snps 0 2>gpurun_out/err.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value']), d['stages_ms'], round(d['roofline']['frac'],4))
"
timeout 200 python tools/bench_variants.py 2>&1 | tail -3
