set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest5.log
timeout 900 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench5_c5.log 2> gpurun_out/r2_bench5_c5.err; echo "rc=$?" >> gpurun_out/r2_bench5_c5.err
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench5.log 2> gpurun_out/r2_bench5.err; echo "rc=$?" >> gpurun_out/r2_bench5.err
