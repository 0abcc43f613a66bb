python -m pytest tests -m gpu -q -x -k "parity or e2e or fullsize" 2>&1 | tail -3
for f in 1 0; do
for nu2 in 5 4 3; do
SB200_SCAN_JSPLIT_FIXED=$f python bench.py --workload cfg5 --nu2 $nu2 --steps 20 --warmup 3 --no-cpu-baseline --clients 0 --sustained-s 0 --tc-batch 0 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('fixed=$f nu2=$nu2 scan ms', round(d['stages_ms']['first_dim_scan'],4), 'GB/s', round(d['roofline']['achieved']), 'frac', round(d['roofline']['frac'],3), 'verified', d['verified']['decoded_equal_planted'])"
done; done
