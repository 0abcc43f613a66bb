#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pack.py -m gpu -q -x 2>&1 | tail -4
run() { # label workload env...
  local label=$1; local wl=$2; shift 2
  env "$@" timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --clients 0 --sustained-s 0 --tc-batch 0 --workloads "" 2>gpurun_out/t_err.log | python -c "
import json,sys
d=json.load(sys.stdin)
print('$label $wl', 'ms', round(d['value'],4), {k: round(v,4) for k,v in d['stages_ms'].items()}, 'frac', round(d['roofline']['frac'],4), [round(x,4) for x in d['roofline']['scan_ms_min_med_max']], 'verified', d['verified']['decoded_equal_planted'], d['roofline']['kernel'])" || tail -3 gpurun_out/t_err.log
}
for n in 7 6; do run shaped_nu2_$n "cfg3 --nu2 $n" X=1; run generic_nu2_$n "cfg3 --nu2 $n" SB200_PACK_SCAN_SHAPED=0; done
