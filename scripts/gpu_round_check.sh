set -x
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s4_pytest.log 2>&1; tail -3 gpurun_out/s4_pytest.log
timeout 600 python bench.py > gpurun_out/s4_bench_1gpu.json 2> gpurun_out/s4_bench_1gpu.err; tail -3 gpurun_out/s4_bench_1gpu.err; python -c "
import json; d=json.load(open('gpurun_out/s4_bench_1gpu.json')); print(d['value'], d['e2e'], d['roofline']['frac'], d['clocks']); print(json.dumps(d.get('pipelined'), indent=1))"
