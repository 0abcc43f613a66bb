#!/bin/bash
# round-2 check on one B200: GPU test suite (incl. the full-size drop-in runs), default bench line, in-graph timeline of a query
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/r2_gpu.txt 2>&1
free -g >> gpurun_out/r2_gpu.txt; nproc >> gpurun_out/r2_gpu.txt
( time python -m pytest tests -m gpu -q -rs --durations=15 ${PYTEST_ARGS} ) > gpurun_out/r2_pytest.log 2>&1
tail -5 gpurun_out/r2_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
tail -c 600 gpurun_out/r2_bench_1gpu.err; head -c 1500 gpurun_out/r2_bench_1gpu.json
python scripts/trace_query.py cfg1 > gpurun_out/r2_trace_cfg1.md 2> gpurun_out/r2_trace_cfg1.err
head -3 gpurun_out/r2_trace_cfg1.md
