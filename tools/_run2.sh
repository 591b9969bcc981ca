set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log
timeout 900 python tools/march_check.py 10 0.75 hll > gpurun_out/r2_march2.log 2>&1; echo "rc=$?" >> gpurun_out/r2_march2.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench2.log 2> gpurun_out/r2_bench2.err; echo "rc=$?" >> gpurun_out/r2_bench2.err
