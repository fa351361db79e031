timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "tensor_core or properties_at_scale or tiny" > gpurun_out/pytest_tc4.log 2>&1; tail -3 gpurun_out/pytest_tc4.log
SPAGXE_TC_CFG=13 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "tensor_core or properties_at_scale or tiny" > gpurun_out/pytest_tc4b.log 2>&1; tail -3 gpurun_out/pytest_tc4b.log
SPAGXE_TC_CFG=16 timeout 300 python bench.py --no-e2e --steps 2 --warmup 1 --cpu-sample-snps 0 > gpurun_out/trace.log 2> gpurun_out/trace.err
grep -c trace gpurun_out/trace.err
