# scan-shape check: parity of the narrow-shard scan variants, then stage timings of the affected shapes with two query-staging sizes
set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_fullsize.py -q -x ) > gpurun_out/pytest_scan.log 2>&1; tail -4 gpurun_out/pytest_scan.log
SH="cfg1:8,7;cfg1:9,6;cfg1:10,5;cfg1:7,6;cfg1:8,5;cfg1:9,8;cfg1:10,7;cfg1:8,9"
timeout 300 python scripts/cost_model_b200.py measure --shapes "$SH" --out gpurun_out/scan_shapes_32k.json 2> gpurun_out/scan_shapes_32k.err
SB200_SCAN_SMEM=16384 timeout 300 python scripts/cost_model_b200.py measure --shapes "$SH" --out gpurun_out/scan_shapes_16k.json 2> gpurun_out/scan_shapes_16k.err
tail -8 gpurun_out/scan_shapes_32k.err; tail -8 gpurun_out/scan_shapes_16k.err
