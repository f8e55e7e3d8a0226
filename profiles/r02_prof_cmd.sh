set -x
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_traffic_launches.csv python profiles/prof_driver.py --reps 4 > gpurun_out/r02_traffic.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_under_ncu.log 2>&1
bash profiles/ncu_capture.sh r02_S1fused "gsb_jit_geo" 0
bash profiles/ncu_capture.sh r02_S2 "k_sweepw.*T3SymS2" 7
bash profiles/ncu_capture.sh r02_S3 "k_sweepw.*TLast" 3
PROF_CG=5 bash profiles/ncu_capture.sh r02_spmv "k_spmv_reg" 2
ls -la gpurun_out | tail -20
