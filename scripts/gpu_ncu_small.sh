set -x
SB200_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fold_decomp_ntt|k_expand_digits|k_fold_mac|k_expand_accum|k_expand_prep|k_fold_lift' -c 48 -o gpurun_out/s4_small_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --clients 0 --tc-batch 0 > gpurun_out/s4_small_ncu.log 2>&1; tail -2 gpurun_out/s4_small_ncu.log | cut -c1-200
