#!/bin/bash
# round-2 ncu evidence (one B200): launch list of one cfg1 query, --set full captures of the scan kernels at the shapes the verdict asked about
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --clients 0 --tc-batch 0 --sustained-s 0 --workloads ''"
SB200_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --clients 0 --tc-batch 0 --sustained-s 0 --workloads "" > /dev/null 2>&1
cap() { # name kernel-regex bench-args...
  local name=$1 rx=$2; shift 2
  SB200_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -o gpurun_out/r2_$name -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --clients 0 --tc-batch 0 --sustained-s 0 --workloads "" "$@" > /dev/null 2> gpurun_out/r2_$name.err
}
cap scan_cfg1 k_scan_spiral
cap scan_cfg5 k_scan_spiral --workload cfg5
cap scan_shard64_9_5 k_scan_spiral --workload cfg5 --nu2 5
cap scan_pack_cfg4 k_scan_pack --workload cfg4
cap scan_pack_cfg3 k_scan_pack --workload cfg3
ls -la gpurun_out/r2_*.ncu-rep
