set -x
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:'k_(fold|from_ntt|scan_pack|expand|pack|rescale|gadget|simple|reorient|gsw)' -c 400 --csv --log-file gpurun_out/cfg3_launches.csv python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu-baseline --tc-batch 0 > gpurun_out/cfg3_ncu.log 2>&1; tail -2 gpurun_out/cfg3_ncu.log | cut -c1-300
