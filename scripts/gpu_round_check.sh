set -x
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s4_pytest.log 2>&1; tail -3 gpurun_out/s4_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s4_bench_1gpu.json 2> gpurun_out/s4_bench_1gpu.err; tail -3 gpurun_out/s4_bench_1gpu.err; python -c "
import json; d=json.load(open('gpurun_out/s4_bench_1gpu.json')); print(d['value'], d['e2e']['value'], d['stages_ms'], d['roofline']['frac']); t=d.get('pipelined',{}).get('tensor_core_batch'); print(t['ms_per_query_amortised'], t['queries_per_s'])"
for w in cfg3; do
timeout 900 python bench.py --workload $w --no-cpu-baseline --steps 8 > gpurun_out/s4_bench_$w.json 2> gpurun_out/s4_bench_$w.err; tail -3 gpurun_out/s4_bench_$w.err; python -c "
import json; d=json.load(open('gpurun_out/s4_bench_$w.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['stages_ms']); t=d.get('pipelined',{}).get('tensor_core_batch'); print(t['ms_per_query_amortised'], t['queries_per_s'])"
done
SB200_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/s4_launches2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --clients 0 --tc-batch 0 > /dev/null 2>&1
