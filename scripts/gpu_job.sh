mkdir -p gpurun_out
SAN_TIMEOUT=700 bash scripts/sanitize.sh racecheck 2>&1 | tail -4
grep -E "Race|hazard|invalid" gpurun_out/sanitize_racecheck.log | sed 's/0x[0-9a-f]*/X/g' | sort | uniq -c | sort -rn | head -10
( timeout 600 python -m pytest tests/test_icp_gpu.py -x -q -m gpu 2>&1 | tail -3 )
