#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --clients 0 --sustained-s 0 --tc-batch 0 --workloads "" 2>gpurun_out/t_err.log | python -c "
import json,sys
d=json.load(sys.stdin)
print('cfg3 x$N', 'ms', round(d['value'],4), {k: round(v,4) for k,v in d['stages_ms'].items()}, 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],4), 'verified', d['verified']['decoded_equal_planted'], d['verified']['owner_ranks'], d['config']['exchange'][:200])" || tail -5 gpurun_out/t_err.log
