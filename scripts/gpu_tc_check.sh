set -x
( time timeout 600 python -m pytest tests/test_gpu_tc.py -x -q ) > gpurun_out/tc_pytest.log 2>&1; tail -15 gpurun_out/tc_pytest.log
timeout 300 python scripts/bench_tc.py > gpurun_out/tc_bench.log 2>&1; tail -12 gpurun_out/tc_bench.log
