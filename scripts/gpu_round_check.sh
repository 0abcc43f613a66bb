set -x
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s4_pytest.log 2>&1; tail -3 gpurun_out/s4_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s4_bench_1gpu.json 2> gpurun_out/s4_bench_1gpu.err; tail -3 gpurun_out/s4_bench_1gpu.err; python -c "
import json; d=json.load(open('gpurun_out/s4_bench_1gpu.json')); print(d['value'], d['e2e']['value'], d['stages_ms'], d['roofline']['frac']); print(json.dumps(d.get('pipelined',{}).get('tensor_core_batch'), indent=1))"
timeout 900 python bench.py --workload cfg3 --no-cpu-baseline --steps 8 > gpurun_out/s4_bench_cfg3.json 2> gpurun_out/s4_bench_cfg3.err; tail -3 gpurun_out/s4_bench_cfg3.err; python -c "
import json; d=json.load(open('gpurun_out/s4_bench_cfg3.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['stages_ms']); print(json.dumps(d.get('pipelined',{}).get('tensor_core_batch'), indent=1))"
