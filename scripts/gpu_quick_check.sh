# short GPU check: the newest test files, then the default bench (gpurun -- bash scripts/gpu_quick_check.sh [pytest args])
set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest ${@:-tests/test_gpu_wire.py tests/test_gpu_client.py} -q -x ) > gpurun_out/pytest_new.log 2>&1; tail -6 gpurun_out/pytest_new.log
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -3 gpurun_out/bench_1gpu.err; cut -c1-1800 gpurun_out/bench_1gpu.json
