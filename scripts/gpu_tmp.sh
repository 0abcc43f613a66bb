python -m pytest tests -m gpu -q -x -k "dropin and not full_size" 2>&1 | tail -15
python -m pytest tests -m gpu -q -x -k "resident" -s 2>&1 | grep -E "harness timers|passed|failed"
python bench.py --steps 30 --warmup 3 --workloads "" --no-cpu-baseline --clients 0 --sustained-s 0 2>&1 | tail -5 | cut -c1-600
