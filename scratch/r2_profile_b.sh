# Round-2 profile refresh (one GPU): op sweep (BASELINE configs[3]), ncu launch list of the headline workload,
# ncu --set full of the top kernels (exported to raw CSV on the box), compute-sanitizer over the op-level tests.
set -x
python scratch/op_sweep.py > gpurun_out/r2_op_sweep.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_sem24k.csv python bench.py --workloads sem24k --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2b_ncu.log 2>&1
python scratch/agg_launches.py gpurun_out/r2_launches_sem24k.csv 40 > gpurun_out/r2_launches_sem24k_summary.txt
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'knn_tc_kernel|fps_pruned|knn3_grid_kernel|knn_tiled' -c 12 -f -o /tmp/prof_geo python bench.py --workloads sem24k --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2b_ncu_geo.log 2>&1
ncu -i /tmp/prof_geo.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_geo_raw.csv 2>/dev/null
for tool in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 86 python -m pytest tests/test_gpu_ops.py tests/test_gpu_knn_grid.py tests/test_gpu_fps_pruned.py tests/test_gpu_knn_tc.py -m gpu -q -x -k "not timing and not large and not full_depth and not 24k and not 24000 and not 24576 and not 13000 and not 12288 and not 12000" > gpurun_out/r2_sanitizer_$tool.log 2>&1; echo "rc=$?" >> gpurun_out/r2_sanitizer_$tool.log
done
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 86 python -m pytest tests/test_gpu_ops.py tests/test_gpu_knn_grid.py tests/test_gpu_fps_pruned.py -m gpu -q -x -k "(fps or knn or transition or gather or interpolate) and not timing and not large and not full_depth and not 24000 and not 24576 and not 13000 and not 12288 and not 12000 and not 5000 and not 4096" > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2_sanitizer_racecheck.log
du -sh gpurun_out
