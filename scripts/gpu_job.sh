mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_torch_ops_gpu.py tests/test_abi_cpu.py -x -q 2>&1 | tail -12 ) | tee gpurun_out/r02v_tests.log
