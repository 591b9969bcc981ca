set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest10.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench10.log 2> gpurun_out/r2_bench10.err; echo "rc=$?" >> gpurun_out/r2_bench10.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches10.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fast-mode > gpurun_out/r2_ncu10.log 2>&1
