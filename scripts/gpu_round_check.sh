set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s4_pytest.log 2>&1; tail -5 gpurun_out/s4_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s4_smoke.log 2>&1; tail -2 gpurun_out/s4_smoke.log
timeout 600 python bench.py > gpurun_out/s4_bench_1gpu.json 2> gpurun_out/s4_bench_1gpu.err; cut -c1-1500 gpurun_out/s4_bench_1gpu.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s4_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --clients 0 > gpurun_out/s4_ncu_bench.log 2>&1; tail -2 gpurun_out/s4_ncu_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_spiral -s 1 -c 2 -o gpurun_out/s4_scan_full python bench.py --steps 2 --warmup 1 --no-cpu-baseline --clients 0 > gpurun_out/s4_ncu_full.log 2>&1; tail -2 gpurun_out/s4_ncu_full.log
