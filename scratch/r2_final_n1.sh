# Final N = 1 measurement cycle of the round (tests, smoke, both bench arms, launch lists, ncu full of knn_tc, GEMM shapes)
python -m pytest tests -m gpu -q > gpurun_out/r4_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4_gputests.log; tail -3 gpurun_out/r4_gputests.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_n1.json 2>> gpurun_out/r2_bench_n1.err; echo "ref rc=$?"
for wl in sem24k partseg2048; do
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_$wl.csv python bench.py --workloads $wl --ncu-workload $wl --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2c_ncu_$wl.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -c 1 -s 1 -f -o /tmp/prof_knntc python scratch/knn_tc_one.py > gpurun_out/r2c_ncu_knntc.log 2>&1
ncu -i /tmp/prof_knntc.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_knn_tc_raw.csv 2>/dev/null
python scratch/gemm_shapes.py sem24k > gpurun_out/r2_gemm_shapes_sem24k.txt 2>&1
rm -f gpurun_out/r2_ops_by_entry.txt; python bench.py --no-cpu-baseline --steps 3 --profile-ops gpurun_out/r2_ops_by_entry.txt > /dev/null 2>&1
python scratch/knn_tc_trace.py > gpurun_out/r2_knn_tc_timeline.txt 2>&1
