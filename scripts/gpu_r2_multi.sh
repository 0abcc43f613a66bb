#!/bin/bash
# round-2 multi-GPU check: bash scripts/gpu_r2_multi.sh N  (run under `gpurun --gpus N`)
N=${1:-2}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"; }
run --steps 20 --warmup 3 --workloads "cfg5,cfg3" > gpurun_out/m_bench_${N}gpu.json 2> gpurun_out/m_bench_${N}gpu.err
tail -c 600 gpurun_out/m_bench_${N}gpu.err
run --workload cfg4 --steps 20 --warmup 3 --sustained-s 0 > gpurun_out/m_bench_cfg4_${N}gpu.json 2> gpurun_out/m_bench_cfg4_${N}gpu.err
tail -c 300 gpurun_out/m_bench_cfg4_${N}gpu.err
python - <<P
import json
for f in ("gpurun_out/m_bench_${N}gpu.json", "gpurun_out/m_bench_cfg4_${N}gpu.json"):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "failed:", e); continue
    print(f, "ms/query", round(d["value"], 4), {k: round(v, 4) for k, v in d["stages_ms"].items()}, "frac", round(d["roofline"]["frac"], 3), "e2e", round(d["e2e"]["value"], 4),
          "sustained", (d.get("sustained") or {}).get("value"), "verified", d["verified"] and (d["verified"]["decoded_equal_planted"], d["verified"]["owner_ranks"]))
    for w, v in d.get("workloads", {}).items():
        print("  ", w, round(v["value"], 4), {k: round(x, 4) for k, x in v["stages_ms"].items()}, "frac", round(v["roofline"]["frac"], 3), "e2e", round(v["e2e"]["value"], 4),
              "verified", v["verified"] and (v["verified"]["decoded_equal_planted"], v["verified"]["owner_ranks"]))
P
