# Round profile refresh (run under gpurun, one GPU).  Keeps gpurun_out/ small: .ncu-rep files are exported to raw CSV
# on the box and not copied back.
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final4_n1.json 2> gpurun_out/bench_final4_n1.err
python bench_workloads.py cls1024 > gpurun_out/wl_cls_n1.json 2>/dev/null
python bench_workloads.py sem24k > gpurun_out/wl_sem_n1.json 2>/dev/null
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1_i.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_i.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:linear_3xtf32 -c 4 -f -o /tmp/prof_gemm python scratch/gemm_one.py 65536 64 64 > gpurun_out/ncu_gemm_c.log 2>&1
ncu -i /tmp/prof_gemm.ncu-rep --page raw --csv > gpurun_out/ncu_full_gemm_c_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:'knn|fps|col_reduce|bn_|attn_|transition_g|transition_b|gather_k' -c 22 -f -o /tmp/prof_misc python scratch/kernels_one.py > gpurun_out/ncu_misc_c.log 2>&1
ncu -i /tmp/prof_misc.ncu-rep --page raw --csv > gpurun_out/ncu_full_misc_c_raw.csv 2>/dev/null
python scratch/gemm_shapes.py > gpurun_out/gemm_shapes_e.txt 2>&1
du -sh gpurun_out
