# Round profile refresh (run under gpurun, one GPU).  Keeps gpurun_out/ small: .ncu-rep files are exported to raw CSV
# and the big one is deleted on the box.
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final2_n1.json 2> gpurun_out/bench_final2_n1.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1_g.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:linear_3xtf32 -c 4 -f -o gpurun_out/prof_gemm_r1b python scratch/gemm_one.py 65536 64 64 > gpurun_out/ncu_gemm_b.log 2>&1
ncu -i gpurun_out/prof_gemm_r1b.ncu-rep --page raw --csv > gpurun_out/ncu_full_gemm_b_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:'knn|fps|col_reduce|bn_|attn_|transition_g|transition_b|gather_k' -c 22 -f -o /tmp/prof_misc_r1b python scratch/kernels_one.py > gpurun_out/ncu_misc_b.log 2>&1
ncu -i /tmp/prof_misc_r1b.ncu-rep --page raw --csv > gpurun_out/ncu_full_misc_b_raw.csv 2>/dev/null
du -sh gpurun_out
