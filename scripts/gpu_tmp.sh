#!/bin/bash
mkdir -p gpurun_out
( time python bench.py ) > gpurun_out/f_bench_1gpu.json 2> gpurun_out/f_bench_1gpu.err
tail -c 400 gpurun_out/f_bench_1gpu.err
( time python bench.py --impl reference --steps 2 --warmup 0 ) > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
tail -c 300 gpurun_out/f_bench_ref.err; head -c 600 gpurun_out/f_bench_ref.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
