# Round-2 first profile: full-size parity file, per-entry-point device-time tables of every workload, ncu launch list of
# the headline workload's timed region (one eager step).
set -x
python -m pytest tests/test_gpu_fullsize.py -m gpu -q > gpurun_out/r2d_fullsize.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_fullsize.log
rm -f gpurun_out/r2_ops_by_entry.txt
python bench.py --no-cpu-baseline --steps 5 --profile-ops gpurun_out/r2_ops_by_entry.txt > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_sem24k.csv python bench.py --workloads sem24k --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2d_ncu.log 2>&1
python scratch/agg_launches.py gpurun_out/r2_launches_sem24k.csv 40 > gpurun_out/r2_launches_sem24k_summary.txt
