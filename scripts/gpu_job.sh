( timeout 600 python -m pytest tests/test_icp_gpu.py -x -q -m gpu -k "adversarial or large_tiles or boundaries" 2>&1 | tail -6 )
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
