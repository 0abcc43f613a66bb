# What a round-end check runs on the GPU box (through `gpurun -- bash scripts/gpu_round_check.sh`): the GPU parity suite,
# smoke(), the default bench, and the ncu launch list of one query; everything lands in gpurun_out/.
set -x
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -3 gpurun_out/bench_1gpu.err; cut -c1-1200 gpurun_out/bench_1gpu.json
SB200_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --clients 0 --tc-batch 0 > /dev/null 2>&1
