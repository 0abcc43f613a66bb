set -x
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s4_pytest.log 2>&1; tail -3 gpurun_out/s4_pytest.log
timeout 600 python bench.py > gpurun_out/s4_bench_1gpu.json 2> gpurun_out/s4_bench_1gpu.err; tail -3 gpurun_out/s4_bench_1gpu.err; python -c "
import json; d=json.load(open('gpurun_out/s4_bench_1gpu.json')); print(d['value'], d['e2e']['value'], d['stages_ms'], d['roofline']['frac'], d['clocks'], d['cpu_baseline']); print(json.dumps(d.get('pipelined',{}).get('tensor_core_batch'), indent=1))"
for w in cfg3 cfg5; do
timeout 900 python bench.py --workload $w --no-cpu-baseline --steps 8 > gpurun_out/s4_bench_$w.json 2> gpurun_out/s4_bench_$w.err; tail -3 gpurun_out/s4_bench_$w.err; python -c "
import json; d=json.load(open('gpurun_out/s4_bench_$w.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['stages_ms'], d['clocks']); print(json.dumps(d.get('pipelined',{}).get('tensor_core_batch'), indent=1))"
done
timeout 600 python bench.py --workload cfg4 --no-cpu-baseline --steps 8 > gpurun_out/s4_bench_cfg4.json 2> gpurun_out/s4_bench_cfg4.err; cut -c1-900 gpurun_out/s4_bench_cfg4.json
