set -x
( time timeout 900 python -m pytest tests/test_gpu_tc.py -x -q ) > gpurun_out/tc_pytest.log 2>&1; tail -12 gpurun_out/tc_pytest.log
TC_COUNTS=1,8,16 TC_ITERS=10 timeout 300 python scripts/bench_tc.py > gpurun_out/tc_bench.log 2>&1; tail -4 gpurun_out/tc_bench.log | cut -c1-400
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/tc_bench_cfg1.json 2> gpurun_out/tc_bench_cfg1.err; tail -3 gpurun_out/tc_bench_cfg1.err; python -c "
import json; d=json.load(open('gpurun_out/tc_bench_cfg1.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac']); print(json.dumps(d.get('pipelined',{}).get('tensor_core_batch'), indent=1))"
timeout 900 python bench.py --workload cfg3 --no-cpu-baseline --steps 8 > gpurun_out/tc_bench_cfg3.json 2> gpurun_out/tc_bench_cfg3.err; tail -3 gpurun_out/tc_bench_cfg3.err; python -c "
import json; d=json.load(open('gpurun_out/tc_bench_cfg3.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['stages_ms']); print(json.dumps(d.get('pipelined',{}).get('tensor_core_batch'), indent=1))"
nvidia-smi --query-gpu=memory.used --format=csv
