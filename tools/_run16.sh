mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_fused_gpu.py tests/test_c4_scale_gpu.py tests/test_rae2822_gpu.py -m gpu -q > gpurun_out/r2_pytest16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest16.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench16.log 2> gpurun_out/r2_bench16.err; echo "rc=$?" >> gpurun_out/r2_bench16.err
