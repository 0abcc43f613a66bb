# What a round-end check runs on the GPU box (through `gpurun -- bash scripts/gpu_round_check.sh`): the GPU parity suite,
# smoke(), the default bench, the shape sweep, the ncu launch list of one query and full captures of the scan kernel
# (cfg1 and the 8 GiB cfg5 shape) plus the wire-query ingest kernel; everything lands in gpurun_out/.
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -3 gpurun_out/bench_1gpu.err; cut -c1-1200 gpurun_out/bench_1gpu.json
timeout 600 python scripts/cost_model_b200.py measure --out gpurun_out/shape_sweep.json 2> gpurun_out/shape_sweep.err; tail -2 gpurun_out/shape_sweep.err
SB200_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --clients 0 --tc-batch 0 > /dev/null 2>&1
SB200_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_spiral -s 2 -c 2 -o gpurun_out/scan_full -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --clients 0 --tc-batch 0 > /dev/null 2>&1
SB200_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_spiral -s 2 -c 2 -o gpurun_out/scan_full_cfg5 -f python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu-baseline --clients 0 --tc-batch 0 > /dev/null 2>&1
SB200_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_query_from_wire -c 2 -o gpurun_out/wire_full -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --clients 0 --tc-batch 0 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
