set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2_gpu.txt; nproc >> gpurun_out/r2_gpu.txt; free -g >> gpurun_out/r2_gpu.txt; df -h /tmp >> gpurun_out/r2_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
timeout 600 python -m pytest tests/test_rae2822_gpu.py tests/test_fused_gpu.py -m gpu -q > gpurun_out/r2_pytest_rae_again.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_rae_again.log
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/r2_memcheck.log python tools/sanitize_paths.py > gpurun_out/r2_memcheck.out 2>&1; echo "rc=$?" >> gpurun_out/r2_memcheck.out
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r2_racecheck.log python tools/sanitize_paths.py > gpurun_out/r2_racecheck.out 2>&1; echo "rc=$?" >> gpurun_out/r2_racecheck.out
timeout 900 compute-sanitizer --tool initcheck --log-file gpurun_out/r2_initcheck.log python tools/sanitize_paths.py > gpurun_out/r2_initcheck.out 2>&1; echo "rc=$?" >> gpurun_out/r2_initcheck.out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench1.log 2> gpurun_out/r2_bench1.err; echo "rc=$?" >> gpurun_out/r2_bench1.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench1_ref.log 2> gpurun_out/r2_bench1_ref.err; echo "rc=$?" >> gpurun_out/r2_bench1_ref.err
