#!/bin/bash
# odd-chain CTA-slot experiment: headline ms/query for several slot limits (0 = unlimited FIFO)
mkdir -p gpurun_out
for s in 0 1 2 3; do
  SB200_ODD_SLOTS=$s python bench.py --steps 30 --warmup 3 --workloads "" --no-cpu-baseline --clients 0 --sustained-s 0 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('slots $s ms/query', round(d['value'],4), {k: round(v,4) for k,v in d['stages_ms'].items()}, 'e2e', round(d['e2e']['value'],4), 'verified', d['verified']['decoded_equal_planted'])"
done
SB200_ODD_SLOTS=2 python scripts/trace_query.py cfg1 > gpurun_out/q_trace_cfg1_slots2.md 2>/dev/null
python -m pytest tests -m gpu -q -x -k "sharded_expansion or pack_client or all_gpu_pack" 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --workloads "cfg3" --no-cpu-baseline --clients 0 --sustained-s 0 --tc-batch 0 2> gpurun_out/q_bench_cfg3.err | python -c "
import json,sys
d=json.load(sys.stdin)
for w,v in d.get('workloads',{}).items(): print(w, round(v['value'],3), v['stages_ms'], round(v['roofline']['frac'],3), 'e2e', round(v['e2e']['value'],3), 'verified', v['verified'])
"
tail -3 gpurun_out/q_bench_cfg3.err
python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline --clients 0 --sustained-s 0 --tc-batch 0 2> gpurun_out/q_bench_cfg4.err | python -c "
import json,sys
v=json.load(sys.stdin)
print('cfg4', round(v['value'],3), v['stages_ms'], round(v['roofline']['frac'],3), 'e2e', round(v['e2e']['value'],3), 'verified', v['verified'])
"
tail -3 gpurun_out/q_bench_cfg4.err
