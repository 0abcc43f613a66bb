set -x
SB200_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/s4_launches3.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --clients 0 --tc-batch 0 > /dev/null 2>&1
