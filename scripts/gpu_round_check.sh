set -x
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s4_pytest.log 2>&1; tail -3 gpurun_out/s4_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s4_bench_1gpu.json 2> gpurun_out/s4_bench_1gpu.err; tail -3 gpurun_out/s4_bench_1gpu.err; python -c "
import json; d=json.load(open('gpurun_out/s4_bench_1gpu.json')); print(d['value'], d['e2e']['value'], d['stages_ms'], d['roofline']['frac']); t=d.get('pipelined',{}).get('tensor_core_batch'); print(t['ms_per_query_amortised'], t['queries_per_s'])"
