set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused_gpu.py tests/test_c4_scale_gpu.py -m gpu -x -q > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log
timeout 900 python tools/march_check.py 10 0.75 hll --analytic > gpurun_out/r2_march3.log 2>&1; echo "rc=$?" >> gpurun_out/r2_march3.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --analytic-sphere > gpurun_out/r2_ncu3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_march_flux -c 3 -o gpurun_out/r2_prof_march3 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --analytic-sphere > gpurun_out/r2_ncu3b.log 2>&1
