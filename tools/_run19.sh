mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches19_c5.csv python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu19.log 2>&1
