set -x
SB200_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fold_decomp_ntt|k_fold_mac|k_fold_lift|k_from_ntt$' -c 8 -o gpurun_out/s4_cfg3_fold python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu-baseline --tc-batch 0 > gpurun_out/s4_cfg3_fold.log 2>&1; tail -2 gpurun_out/s4_cfg3_fold.log | cut -c1-200
