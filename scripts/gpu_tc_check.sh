set -x
( time timeout 600 python -m pytest tests/test_gpu_tc.py -x -q ) > gpurun_out/tc_pytest.log 2>&1; tail -5 gpurun_out/tc_pytest.log
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv -lms 100 > gpurun_out/tc_clocks.csv &
SMI=$!
TC_COUNTS=1,4,8,10,12,16 TC_ITERS=20 timeout 300 python scripts/bench_tc.py > gpurun_out/tc_bench_sustained.log 2>&1; tail -7 gpurun_out/tc_bench_sustained.log
kill $SMI
sort gpurun_out/tc_clocks.csv | uniq -c | sort -rn | head -8
TC_COUNTS=16 TC_ITERS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_tc -s 2 -c 1 -o gpurun_out/tc_scan_full python scripts/bench_tc.py > gpurun_out/tc_ncu.log 2>&1; tail -3 gpurun_out/tc_ncu.log
