# scan-shape check: stage timings of the shapes the scan tiling depends on, under the tuning knobs of launch_scan_spiral
set -x
mkdir -p gpurun_out
SH="cfg1:10,5;cfg1:8,5;cfg1:9,5;cfg1:8,7"
for v in 0 1; do
  SB200_SCAN_T64=$v timeout 300 python scripts/cost_model_b200.py measure --shapes "$SH" --out gpurun_out/scan_shapes_t64_$v.json 2> gpurun_out/scan_shapes_t64_$v.err
done
timeout 600 python bench.py --workload cfg5 > gpurun_out/bench_cfg5_1gpu.json 2> gpurun_out/bench_cfg5_1gpu.err; tail -2 gpurun_out/bench_cfg5_1gpu.err
