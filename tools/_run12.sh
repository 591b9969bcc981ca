mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_fused_gpu.py tests/test_c4_scale_gpu.py tests/test_rans_gpu.py -m gpu -q > gpurun_out/r2_pytest12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest12.log
timeout 900 python tools/march_check.py 10 0.75 hll --analytic > gpurun_out/r2_march12.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches12.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fast-mode --analytic-sphere > gpurun_out/r2_ncu12.log 2>&1
