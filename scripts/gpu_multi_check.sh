# N-GPU check (gpurun --gpus N -- bash scripts/gpu_multi_check.sh N): the N-GPU weak-scaling bench (cfg1 shard per GPU) and,
# with a second argument, one more workload in its strong-scaling form
set -x
N=${1:-2}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 "$@"; }
run > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; tail -2 gpurun_out/bench_${N}gpu.err; cut -c1-400 gpurun_out/bench_${N}gpu.json
if [ -n "$2" ]; then run --workload $2 > gpurun_out/bench_$2_${N}gpu.json 2> gpurun_out/bench_$2_${N}gpu.err; cut -c1-400 gpurun_out/bench_$2_${N}gpu.json; fi
