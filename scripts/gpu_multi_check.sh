# N-GPU bench of the default workload, launched the way the driver does: bash scripts/gpu_multi_check.sh N  (under gpurun --gpus N)
set -x
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err || tail -8 gpurun_out/bench_${N}gpu.err
cut -c1-1600 gpurun_out/bench_${N}gpu.json
