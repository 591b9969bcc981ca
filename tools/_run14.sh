mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest14.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench14.log 2> gpurun_out/r2_bench14.err; echo "rc=$?" >> gpurun_out/r2_bench14.err
