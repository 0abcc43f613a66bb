#!/bin/bash
# tuning helper: first-dimension scan kernel variants at cfg1 (run on the GPU box)
for v in 0 4 0 4; do
  SB200_SCAN_VARIANT=$v python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('variant $v', 'scan_ms', d['roofline']['scan_ms_min_med_max'], 'frac', round(d['roofline']['frac'],4), 'ms/query', round(d['value'],4))"
done
