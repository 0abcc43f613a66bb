for i in 1 2; do python -m pytest tests -m gpu -q -k "sharded_expansion or pack_peer" 2>&1 | grep -E "AssertionError|passed|failed" | head -3; done
CUDA_MODULE_LOADING=LAZY python -m pytest tests -m gpu -q -k "sharded_expansion and 6-3-2" 2>&1 | grep -E "AssertionError|passed|failed" | head -3
