mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rans_gpu.py tests/test_fused_gpu.py -m gpu -q > gpurun_out/r2_pytest18.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest18.log
timeout 900 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench18_c5.log 2> gpurun_out/r2_bench18_c5.err; echo "rc=$?" >> gpurun_out/r2_bench18_c5.err
