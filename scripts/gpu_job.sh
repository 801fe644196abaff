mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_chamfer_gpu.py tests/test_torch_ops_gpu.py tests/test_integration_gpu.py -x -q -m gpu 2>&1 | tail -4 )
timeout 600 python scripts/bench_chamfer.py r02 2>&1 | tail -6
cp profiles/r02_chamfer.md gpurun_out/
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'chamfer_fwd|chamfer_bwd' -s 8 -c 2 -f -o gpurun_out/r02_prof_chamfer python scripts/profile_chamfer.py > gpurun_out/r02_prof_chamfer.log 2>&1
