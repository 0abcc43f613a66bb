# scan-shape check: parity suites that reach every scan tiling, then stage timings of narrow shapes with / without the j-split kernel
set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_e2e.py tests/test_gpu_cli.py -q -x ) > gpurun_out/pytest_scan.log 2>&1; tail -4 gpurun_out/pytest_scan.log
SH="cfg1:10,5;cfg1:8,5;cfg1:9,4;cfg1:10,3"
for v in 1 0; do
  SB200_SCAN_JSPLIT=$v timeout 300 python scripts/cost_model_b200.py measure --shapes "$SH" --out gpurun_out/scan_shapes_jsplit_$v.json 2> gpurun_out/scan_shapes_jsplit_$v.err
done
