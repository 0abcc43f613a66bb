set -x
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -3
timeout 300 python scripts/bench_tc.py 9 8 2>&1 | tail -3 | cut -c1-330
