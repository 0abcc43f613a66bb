#!/bin/bash
mkdir -p gpurun_out
for c in 0 370 740 1480; do
  SB200_ODD_CHUNK=$c python bench.py --steps 30 --warmup 3 --workloads "" --no-cpu-baseline --clients 0 --sustained-s 0 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('chunk $c ms/query', round(d['value'],4), {k: round(v,4) for k,v in d['stages_ms'].items()}, 'e2e', round(d['e2e']['value'],4), 'verified', d['verified']['decoded_equal_planted'])"
done
python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline --clients 0 --sustained-s 0 2> gpurun_out/q_bench_cfg2.err | python -c "
import json,sys
v=json.load(sys.stdin)
print('cfg2', round(v['value'],4), v['stages_ms'], round(v['roofline']['frac'],3), 'e2e', round(v['e2e']['value'],4), 'verified', v['verified']['decoded_equal_planted'], v['config']['workload'])
"
tail -3 gpurun_out/q_bench_cfg2.err
python -m pytest tests -m gpu -q -x -k "pack_peer_exchange or sharded_expansion or implicit" 2>&1 | tail -3
