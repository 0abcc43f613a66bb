#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_tc.py tests/test_gpu_e2e.py tests/test_gpu_wire.py tests/test_gpu_pack.py -m gpu -q -x 2>&1 | tail -5
