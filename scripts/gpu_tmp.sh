python -m pytest tests -m gpu -q -x -k "not full_size and not dropin and not cli and not wire and not tc and not client" 2>&1 | tail -3
python bench.py --steps 30 --warmup 3 --workloads "" --no-cpu-baseline --clients 0 --sustained-s 0 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('ms/query', round(d['value'],4), {k: round(v,4) for k,v in d['stages_ms'].items()}, 'e2e', round(d['e2e']['value'],4), 'verified', d['verified']['decoded_equal_planted'])"
SB200_PROFILE_SKIP_ODD_CHAIN=1 python scripts/trace_query.py cfg1 > gpurun_out/q_trace_marks.md 2>/dev/null
head -24 gpurun_out/q_trace_marks.md | cut -c1-110
python scripts/trace_query.py cfg1 > gpurun_out/q_trace_cfg1.md 2>/dev/null
