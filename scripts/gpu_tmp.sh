#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pack.py -m gpu -q -x 2>&1 | tail -4
run() { # label workload env...
  local label=$1; local wl=$2; shift 2
  env "$@" python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --clients 0 --sustained-s 0 --tc-batch 0 --workloads "" 2>gpurun_out/t_err.log | python -c "
import json,sys
d=json.load(sys.stdin)
print('$label $wl', 'ms', round(d['value'],4), {k: round(v,4) for k,v in d['stages_ms'].items()}, 'frac', round(d['roofline']['frac'],4), d['roofline']['scan_ms_min_med_max'], 'e2e', round(d['e2e']['value'],4), 'verified', d['verified']['decoded_equal_planted'], d['clocks']['reasons'])" || tail -3 gpurun_out/t_err.log
}
run split2 cfg4 X=1
run split1 cfg4 SB200_PACK_SCAN_SPLIT=1
