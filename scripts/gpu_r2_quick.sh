#!/bin/bash
# quick round-2 iteration on one B200: GPU tests without the minute-long full-size drop-in runs, headline bench, in-graph timeline
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x -k "not full_size" ${PYTEST_ARGS} ) > gpurun_out/q_pytest.log 2>&1
tail -4 gpurun_out/q_pytest.log
python bench.py --steps 20 --warmup 3 --workloads "${WORKLOADS}" --no-cpu-baseline --clients 0 > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
tail -c 400 gpurun_out/q_bench.err; python - <<'P'
import json
try:
    d = json.load(open("gpurun_out/q_bench.json"))
    print("ms/query", d["value"], d["stages_ms"], "e2e", d["e2e"]["value"], "sustained", (d.get("sustained") or {}).get("value"), "verified", d["verified"]["decoded_equal_planted"])
    for w, v in d.get("workloads", {}).items():
        print(w, v["value"], v["stages_ms"], v["roofline"]["frac"], v["e2e"]["value"], v["verified"] and v["verified"]["decoded_equal_planted"])
except Exception as e:
    print("bench failed:", e)
P
python scripts/trace_query.py cfg1 > gpurun_out/q_trace_cfg1.md 2> gpurun_out/q_trace.err
