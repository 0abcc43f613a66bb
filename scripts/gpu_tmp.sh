#!/bin/bash
mkdir -p gpurun_out
for x in 0 1; do for wl in cfg5 cfg1; do
SB200_EXPERIMENT_CONTIG_ROWS=$x SB200_BENCH_SKIP_VERIFY=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --clients 0 --sustained-s 0 --tc-batch 0 --workloads "" 2>gpurun_out/t_err.log | python -c "
import json,sys
d=json.load(sys.stdin)
print('contig=$x $wl', 'ms', round(d['value'],4), {k: round(v,4) for k,v in d['stages_ms'].items()}, 'verified', d['verified']['decoded_equal_planted'])" || tail -5 gpurun_out/t_err.log
done; done
