N=2
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"; }
run --workload cfg4 --steps 20 --warmup 3 --sustained-s 0 > gpurun_out/m_bench_cfg4_${N}gpu.json 2> gpurun_out/m_bench_cfg4_${N}gpu.err
grep -v "^\*\|W1017\|OMP" gpurun_out/m_bench_cfg4_${N}gpu.err | tail -5
python - <<P
import json
d = json.load(open("gpurun_out/m_bench_cfg4_2gpu.json"))
print("cfg4 N=2 ms/query", round(d["value"], 4), d["stages_ms"], "frac", round(d["roofline"]["frac"], 3), "e2e", round(d["e2e"]["value"], 4), d["e2e"], "verified", d["verified"]["decoded_equal_planted"], d["config"]["exchange"][:80])
P
