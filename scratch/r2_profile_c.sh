# Round-2 final profile set (one GPU).  Outputs under gpurun_out/, copied into profiles/ by hand.
set -x
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_n1.json 2>> gpurun_out/r2_bench_n1.err
rm -f gpurun_out/r2_ops_by_entry.txt
python bench.py --no-cpu-baseline --steps 3 --profile-ops gpurun_out/r2_ops_by_entry.txt > /dev/null 2>&1
python scratch/op_sweep.py > gpurun_out/r2_op_sweep.txt 2>&1
python scratch/transition_parts.py > gpurun_out/r2_transition_parts.txt 2>&1
python scratch/gemm_shapes.py partseg2048 > gpurun_out/r2_gemm_shapes_partseg.txt 2>&1
python scratch/gemm_shapes.py sem24k > gpurun_out/r2_gemm_shapes_sem24k.txt 2>&1
for wl in sem24k partseg2048; do
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_$wl.csv python bench.py --workloads $wl --ncu-workload $wl --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2c_ncu_$wl.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -c 1 -s 1 -f -o /tmp/prof_knntc python scratch/knn_tc_one.py > gpurun_out/r2c_ncu_knntc.log 2>&1
ncu -i /tmp/prof_knntc.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_knn_tc_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:linear_3xtf32 -c 4 -f -o /tmp/prof_gemm python scratch/gemm_one.py 65536 64 64 > gpurun_out/r2c_ncu_gemm.log 2>&1
ncu -i /tmp/prof_gemm.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_linear_3xtf32_raw.csv 2>/dev/null
ncu --set full --clock-control none --profile-from-start off -k regex:'fps_pruned|knn3_grid_kernel|knn_tiled|linear_bf16|transition_gather_group|knn_few' -c 10 -f -o /tmp/prof_misc python bench.py --workloads sem24k --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2c_ncu_misc.log 2>&1
ncu -i /tmp/prof_misc.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_misc_raw.csv 2>/dev/null
du -sh gpurun_out
