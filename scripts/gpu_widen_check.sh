# GPU check of the rows either side of the hot path (wire formats, record / snapshot I/O, GPU client) followed by the whole
# GPU suite, the shape sweep of scripts/cost_model_b200.py and the default bench: `gpurun -- bash scripts/gpu_widen_check.sh`
set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_wire.py tests/test_gpu_client.py -q -x ) > gpurun_out/pytest_new.log 2>&1; tail -15 gpurun_out/pytest_new.log
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python scripts/cost_model_b200.py measure --out gpurun_out/shape_sweep.json 2> gpurun_out/shape_sweep.err; tail -3 gpurun_out/shape_sweep.err
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -3 gpurun_out/bench_1gpu.err; cut -c1-1500 gpurun_out/bench_1gpu.json
